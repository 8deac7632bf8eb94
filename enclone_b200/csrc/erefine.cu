// erefine.cu — post-join refinement on the GPU (include/enclone_refine.h).
//
// What enclone does with `eq` and `raw_joins` after join_exacts: clonotype chain grouping (define_mat:
// /root/reference/pages/heuristics.html.src:19-41), pure subclonotypes, the doublet / signature / weak-chain filters
// (/root/reference/pages/default_filters.html.src:292-353) and the break-up of clonotypes that are no longer held together
// by shared chains (:315-317).  The reference walks the clonotypes one rayon task each and rebuilds a matrix per clonotype;
// here every clonotype is handled at once in flat arrays:
//   * ONE union-find over all chains of all exact subclonotypes says which chains share a column (raw joins, the
//     third-chain rule, onesies matched by identical V..J through a sort by (clonotype, sequence id));
//   * a second union-find over the exact subclonotypes gives the clonotypes after the break-up;
//   * column numbers come from one sort of the chain orbits by (clonotype, position of the representative) and a scan;
//   * pure subclonotypes are the runs of a sort by a 64-bit hash of (clonotype, columns), verified element by element;
//   * the doublet filter is a search per pure subclonotype over the partners it shares a CDR3 with (two sorted lists:
//     by CDR3 with the largest partners first, and by pure subclonotype), the signature filter is inclusion-exclusion over
//     per-column cell counts, the weak-chain filter per-column cell counts against per-clonotype totals.
// Clause-by-clause provenance ([DOC] / [RECALL]) is in oracle/refine_oracle.c, which tests compare this file against.
// No CPU fallback: without a device erefine_run fails with -3.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include "../../include/enclone_refine.h"

namespace {

thread_local std::string r_err;
int rfail(int code, const std::string &m) { r_err = m; return code; }
#define RCK(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return rfail(-2, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

typedef unsigned long long u64;
#define NONE32 0xFFFFFFFFu
#define RERR_RANGE 1u          // an index of the pack points outside its table
#define RERR_HASH 2u           // two different column signatures with one 64-bit hash

struct RDev {
    uint32_t n, nx, nc; u64 nj;
    const uint32_t *exact_id, *class_of, *jk1, *jk2, *chain_off, *ncells, *origin_rank, *seq_id, *cdr3_id, *order_rank, *seq_len;
    const uint8_t *cols0, *cols1, *left, *seq_arena;
    const uint32_t *off2;          // offsets into the compacted sequence bytes (third chains only)
};
struct RCounters { u64 n_del, triples, n_pure; unsigned int err, pad; };

__device__ __forceinline__ uint32_t uf_find(uint32_t *p, uint32_t x) {
    for (;;) {
        const uint32_t px = __ldcg(&p[x]);
        if (px == x) return x;
        const uint32_t gp = __ldcg(&p[px]);
        if (gp != px) p[x] = gp;
        x = px;
    }
}
__device__ __forceinline__ void uf_union(uint32_t *p, uint32_t a, uint32_t b) {      // the smaller index becomes the root
    for (;;) {
        a = uf_find(p, a); b = uf_find(p, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        if (atomicCAS(&p[a], a, b) == a) return;
    }
}
__device__ __forceinline__ u64 mix64(u64 h, u64 v) {          // splitmix-style step
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 27; h *= 0x94D049BB133111EBull; h ^= h >> 31;
    return h;
}

struct RMax { template <typename T> __host__ __device__ __forceinline__ T operator()(const T &a, const T &b) const { return a > b ? a : b; } };

#define TID (blockIdx.x * (size_t)blockDim.x + threadIdx.x)

// ---- set-up (once) ----
__global__ void k_r_xof(RDev d, uint32_t *xof, RCounters *ctr) {
    const size_t x = TID;
    if (x >= d.nx) return;
    const uint32_t lo = d.chain_off[x], hi = d.chain_off[x + 1];
    if (hi < lo || hi > d.nc || hi - lo > EREFINE_MAX_CHAINS) { atomicOr(&ctr->err, RERR_RANGE); return; }
    for (uint32_t c = lo; c < hi; c++) xof[c] = (uint32_t)x;
}
__global__ void k_r_key_origin(RDev d, const uint32_t *xof, u64 *keys, uint32_t *vals) {
    const size_t c = TID;
    if (c < d.nc) { keys[c] = ((u64)d.origin_rank[xof[c]] << 32) | c; vals[c] = (uint32_t)c; }
}
__global__ void k_r_key_order(RDev d, const uint32_t *pos2, u64 *keys, uint32_t *vals) {
    const size_t c = TID;
    if (c < d.nc) { keys[c] = ((u64)d.order_rank[c] << 32) | pos2[c]; vals[c] = (uint32_t)c; }
}
__global__ void k_r_scatter_pos(const uint32_t *perm, uint32_t n, uint32_t *pos) {
    const size_t i = TID;
    if (i < n) pos[perm[i]] = (uint32_t)i;
}
__global__ void k_r_class_min(RDev d, uint32_t *cmin, RCounters *ctr) {
    const size_t k = TID;
    if (k >= d.n) return;
    const uint32_t x = d.exact_id[k], cl = d.class_of[k];
    if (x >= d.nx || cl >= d.n) { atomicOr(&ctr->err, RERR_RANGE); return; }
    const uint32_t m = d.chain_off[x + 1] - d.chain_off[x];
    if (d.cols0[k] >= m || (d.cols1[k] != 0xFF && d.cols1[k] >= m)) { atomicOr(&ctr->err, RERR_RANGE); return; }
    atomicMin(&cmin[cl], x);
}
__global__ void k_r_label0(RDev d, const uint32_t *cmin, uint32_t *label, uint8_t *alive) {
    const size_t k = TID;
    if (k >= d.n) return;
    const uint32_t x = d.exact_id[k], cl = d.class_of[k];
    if (x >= d.nx || cl >= d.n) return;
    alive[x] = 1;
    atomicMin(&label[x], cmin[cl]);
}
__global__ void k_r_check_joins(RDev d, RCounters *ctr) {
    const size_t j = TID;
    if (j < d.nj && (d.jk1[j] >= d.n || d.jk2[j] >= d.n)) atomicOr(&ctr->err, RERR_RANGE);
}

// ---- one stage ----
__global__ void k_r_iota(uint32_t *p, uint32_t n) { const size_t i = TID; if (i < n) p[i] = (uint32_t)i; }
// S3 (oracle/refine_oracle.c), once per call: the third chains of a raw join between two three-chain exact subclonotypes.  The rule
// does not depend on which exact subclonotypes are still alive, so it is evaluated once per join (third8) and the stages only read
// the answer; the sequences it compares are the only V..J bytes the refinement ever reads, so only those go to the device.
__device__ __forceinline__ bool third_pair(const RDev &d, size_t j, uint32_t *t1, uint32_t *t2) {
    const uint32_t k1 = d.jk1[j], k2 = d.jk2[j];
    const uint32_t x1 = d.exact_id[k1], x2 = d.exact_id[k2];
    const uint32_t o1 = d.chain_off[x1], o2 = d.chain_off[x2];
    const uint8_t a0 = d.cols0[k1], a1 = d.cols1[k1], b0 = d.cols0[k2], b1 = d.cols1[k2];
    if (d.chain_off[x1 + 1] - o1 != 3 || d.chain_off[x2 + 1] - o2 != 3 || a0 >= 3 || a1 >= 3 || a0 == a1 || b0 >= 3 || b1 >= 3 || b0 == b1) return false;
    *t1 = o1 + (3 - a0 - a1); *t2 = o2 + (3 - b0 - b1);
    return d.left[*t1] == d.left[*t2] && d.seq_len[*t1] == d.seq_len[*t2];
}
__global__ void k_r_third_need(RDev d, uint8_t *need) {
    const size_t j = TID;
    uint32_t t1, t2;
    if (j < d.nj && third_pair(d, j, &t1, &t2)) { need[t1] = 1; need[t2] = 1; }
}
__global__ void k_r_need_len(RDev d, const uint8_t *need, uint32_t *len) {
    const size_t c = TID;
    if (c < d.nc) len[c] = need[c] ? d.seq_len[c] : 0u;
}
__global__ void k_r_third(RDev d, int third_diffs, uint8_t *third8) {
    const size_t j = TID;
    if (j >= d.nj) return;
    uint32_t t1, t2;
    uint8_t ok = 0;
    if (third_pair(d, j, &t1, &t2)) {
        const uint8_t *s1 = d.seq_arena + d.off2[t1], *s2 = d.seq_arena + d.off2[t2];
        const uint32_t len = d.seq_len[t1];
        int diffs = 0;
        for (uint32_t i = 0; i < len && diffs <= third_diffs; i++) diffs += s1[i] != s2[i];
        ok = diffs <= third_diffs;
    }
    third8[j] = ok;
}
// S2 + S3 per stage: the raw joins' chain correspondences, and the third chains where k_r_third said yes
__global__ void k_r_joins(RDev d, const uint8_t *alive, const uint8_t *third8, uint32_t *cp) {
    const size_t j = TID;
    if (j >= d.nj) return;
    const uint32_t k1 = d.jk1[j], k2 = d.jk2[j];
    const uint32_t x1 = d.exact_id[k1], x2 = d.exact_id[k2];
    if (!alive[x1] || !alive[x2]) return;
    const uint32_t o1 = d.chain_off[x1], o2 = d.chain_off[x2];
    const uint8_t a0 = d.cols0[k1], a1 = d.cols1[k1], b0 = d.cols0[k2], b1 = d.cols1[k2];
    uf_union(cp, o1 + a0, o2 + b0);
    if (a1 != 0xFF && b1 != 0xFF) uf_union(cp, o1 + a1, o2 + b1);
    if (third8[j]) uf_union(cp, o1 + (3 - a0 - a1), o2 + (3 - b0 - b1));
}
// S4: chains keyed by (clonotype, V..J sequence id)
__global__ void k_r_seq_keys(RDev d, const uint32_t *xof, const uint8_t *alive, const uint32_t *label, u64 *keys, uint32_t *vals) {
    const size_t c = TID;
    if (c >= d.nc) return;
    const uint32_t x = xof[c];
    keys[c] = alive[x] ? (((u64)label[x] << 32) | d.seq_id[c]) : ~0ull;
    vals[c] = (uint32_t)c;
}
__global__ void k_r_heads(const u64 *keys, uint32_t n, u64 mask, uint32_t *head) {       // index of the element if it starts a run of equal (masked) keys, else 0
    const size_t i = TID;
    if (i < n) head[i] = (i == 0 || (keys[i] & mask) != (keys[i - 1] & mask)) ? (uint32_t)i : 0u;
}
// S4 per stage, on the order sorted ONCE by (clonotype the join left, sequence id): a later stage's clonotypes only split the earlier
// ones, and chains that an alive onesie tied together in one stage are in one clonotype in the next, so the runs of the first sort
// serve every stage; only "is there an alive onesie in the run" changes.
__global__ void k_r_onesie_first(RDev d, const uint32_t *perm, const uint32_t *xof, const uint8_t *alive, const uint32_t *run_head, uint32_t *first) {
    const size_t i = TID;
    if (i >= d.nc) return;
    const uint32_t c = perm[i], x = xof[c];
    if (alive[x] && d.chain_off[x + 1] - d.chain_off[x] == 1) atomicMin(&first[run_head[i]], c);
}
__global__ void k_r_onesie_join(RDev d, const uint32_t *perm, const uint32_t *xof, const uint8_t *alive, const uint32_t *run_head, const uint32_t *first, uint32_t *cp) {
    const size_t i = TID;
    if (i >= d.nc) return;
    const uint32_t c = perm[i];
    if (!alive[xof[c]]) return;
    const uint32_t f = first[run_head[i]];
    if (f != NONE32 && f != c) uf_union(cp, c, f);
}
// S6: exact subclonotypes that share a column are one clonotype
__global__ void k_r_link_exacts(RDev d, const uint32_t *xof, const uint8_t *alive, uint32_t *cp, uint32_t *xp) {
    const size_t c = TID;
    if (c >= d.nc || !alive[xof[c]]) return;
    const uint32_t r = uf_find(cp, (uint32_t)c);
    if (r != c) uf_union(xp, xof[c], xof[r]);
}
__global__ void k_r_labels(RDev d, const uint8_t *alive, uint32_t *xp, uint32_t *label) {
    const size_t x = TID;
    if (x < d.nx) label[x] = alive[x] ? uf_find(xp, (uint32_t)x) : NONE32;
}
// S5: representative (first member in (u, m) order) and multiplicity of every chain orbit
__global__ void k_r_orbit_stats(RDev d, const uint8_t *alive, uint32_t *cp, const uint32_t *pos, const uint32_t *pos2, u64 *rep, uint32_t *mm) {
    const size_t x = TID;
    if (x >= d.nx || !alive[x]) return;
    const uint32_t o = d.chain_off[x], m = d.chain_off[x + 1] - o;
    uint32_t r[EREFINE_MAX_CHAINS];
    for (uint32_t a = 0; a < m; a++) r[a] = uf_find(cp, o + a);
    for (uint32_t a = 0; a < m; a++) {
        atomicMin(&rep[r[a]], ((u64)pos2[o + a] << 32) | pos[o + a]);
        uint32_t cnt = 0;
        for (uint32_t b = 0; b < m; b++) cnt += r[b] == r[a];
        if (cnt > 1 || mm[r[a]] == 0) atomicMax(&mm[r[a]], cnt);
    }
}
__global__ void k_r_orbit_keys(RDev d, const uint32_t *xof, const uint8_t *alive, const uint32_t *label, const uint32_t *cp, const u64 *rep, int pb, u64 *keys, uint32_t *vals) {
    const size_t c = TID;
    if (c >= d.nc) return;
    const bool root = alive[xof[c]] && cp[c] == c;
    keys[c] = root ? (((u64)label[xof[c]] << pb) | (rep[c] & 0xFFFFFFFFull)) : ~0ull;       // pb = bits of a chain position
    vals[c] = (uint32_t)c;
}
__global__ void k_r_gather_mm(RDev d, const u64 *keys, const uint32_t *perm, const uint32_t *mm, uint32_t *out) {
    const size_t i = TID;
    if (i < d.nc) out[i] = keys[i] == ~0ull ? 0u : mm[perm[i]];
}
// column base of every orbit = (global exclusive sum of mm) - (the same at the first orbit of its clonotype); the global number
// of a clonotype's first column and its column count go to tables indexed by the clonotype's name (an exact id)
__global__ void k_r_orbit_base(RDev d, const u64 *keys, const uint32_t *perm, const uint32_t *run_head, const uint32_t *gsum, const uint32_t *mms,
                               int pb, uint32_t *base, uint32_t *gbase, uint32_t *ncols_label) {
    const size_t i = TID;
    if (i >= d.nc || keys[i] == ~0ull) return;
    const uint32_t h = run_head[i], lab = (uint32_t)(keys[i] >> pb);
    base[perm[i]] = gsum[i] - gsum[h];
    if (h == i) gbase[lab] = gsum[i];
    atomicAdd(&ncols_label[lab], mms[i]);
}
// columns of every chain; hash of (clonotype, sorted columns) per exact subclonotype
__global__ void k_r_columns(RDev d, const uint8_t *alive, const uint32_t *label, uint32_t *cp, const uint32_t *base, uint32_t *col, u64 *hkeys, uint32_t *hvals) {
    const size_t x = TID;
    if (x >= d.nx) return;
    const uint32_t o = d.chain_off[x], m = d.chain_off[x + 1] - o;
    hvals[x] = (uint32_t)x;
    if (!alive[x]) { for (uint32_t a = 0; a < m; a++) col[o + a] = NONE32; hkeys[x] = ~0ull; return; }
    uint32_t r[EREFINE_MAX_CHAINS], cs[EREFINE_MAX_CHAINS];
    for (uint32_t a = 0; a < m; a++) r[a] = uf_find(cp, o + a);
    for (uint32_t a = 0; a < m; a++) {
        uint32_t before = 0;
        for (uint32_t b = 0; b < a; b++) before += r[b] == r[a];
        cs[a] = base[r[a]] + before;
        col[o + a] = cs[a];
    }
    for (uint32_t a = 1; a < m; a++) { const uint32_t v = cs[a]; uint32_t b = a; while (b > 0 && cs[b - 1] > v) { cs[b] = cs[b - 1]; b--; } cs[b] = v; }
    u64 h = mix64(0x243F6A8885A308D3ull, label[x]);
    for (uint32_t a = 0; a < m; a++) h = mix64(h, cs[a]);
    hkeys[x] = h == ~0ull ? 0 : h;
}
__device__ __forceinline__ bool same_signature(const RDev &d, const uint32_t *label, const uint32_t *col, uint32_t x, uint32_t y) {
    const uint32_t ox = d.chain_off[x], mx = d.chain_off[x + 1] - ox, oy = d.chain_off[y], my = d.chain_off[y + 1] - oy;
    if (label[x] != label[y] || mx != my) return false;
    for (uint32_t a = 0; a < mx; a++) {        // equal as sets (columns of one exact subclonotype are distinct)
        bool f = false;
        for (uint32_t b = 0; b < my; b++) f |= col[ox + a] == col[oy + b];
        if (!f) return false;
    }
    return true;
}
__global__ void k_r_pure_heads(RDev d, const u64 *hk, const uint32_t *hperm, const uint32_t *label, const uint32_t *col, uint32_t *flag, RCounters *ctr) {
    const size_t i = TID;
    if (i >= d.nx) return;
    if (hk[i] == ~0ull) { flag[i] = 0; return; }
    const bool head = i == 0 || hk[i] != hk[i - 1];
    if (!head && !same_signature(d, label, col, hperm[i], hperm[i - 1])) atomicOr(&ctr->err, RERR_HASH);
    flag[i] = head ? 1u : 0u;
}
__global__ void k_r_pure_ids(RDev d, const u64 *hk, const uint32_t *hperm, const uint32_t *incl, uint32_t *pure, u64 *npure, u64 *pure_hash, RCounters *ctr) {
    const size_t i = TID;
    if (i >= d.nx) return;
    const uint32_t x = hperm[i];
    if (hk[i] == ~0ull) { pure[x] = NONE32; return; }
    const uint32_t p = incl[i] - 1;
    pure[x] = p;
    atomicAdd(&npure[p], (u64)d.ncells[x]);
    if (i == 0 || hk[i] != hk[i - 1]) pure_hash[p] = hk[i];      // ascending in p: the signature filter looks pures up by hash
    if (i + 1 == d.nx || hk[i + 1] == ~0ull) ctr->n_pure = p + 1;
}

// ---- F1: doublets ----
__global__ void k_r_content(RDev d, const uint32_t *xof, const uint8_t *alive, const uint32_t *pure, u64 *keys) {
    const size_t c = TID;
    if (c >= d.nc) return;
    const uint32_t x = xof[c];
    keys[c] = alive[x] ? (((u64)d.cdr3_id[c] << 32) | pure[x]) : ~0ull;
}
// by CDR3, the partners with the most cells first: key = cdr3 | ~min(cells, 2^32-1), value = pure
__global__ void k_r_content_by_cells(const u64 *uniq, const uint32_t *n_uniq, const u64 *npure, u64 *keys, uint32_t *vals, u64 *keys_p, uint32_t cap) {
    const size_t i = TID;
    if (i >= cap) return;
    const uint32_t n = *n_uniq;
    const bool live = i < n && uniq[i] != ~0ull;
    if (!live) { keys[i] = ~0ull; vals[i] = NONE32; keys_p[i] = ~0ull; return; }
    const uint32_t cd = (uint32_t)(uniq[i] >> 32), p = (uint32_t)uniq[i];
    const u64 cells = npure[p];
    keys[i] = ((u64)cd << 32) | (0xFFFFFFFFu - (uint32_t)(cells > 0xFFFFFFFFull ? 0xFFFFFFFFull : cells));
    vals[i] = p;
    keys_p[i] = ((u64)p << 32) | cd;
}
__device__ __forceinline__ uint32_t lower_bound64(const u64 *a, uint32_t n, u64 v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ bool pures_share(const u64 *by_p, uint32_t m, uint32_t p1, uint32_t p2) {
    uint32_t i = lower_bound64(by_p, m, (u64)p1 << 32), j = lower_bound64(by_p, m, (u64)p2 << 32);
    while (i < m && j < m && (uint32_t)(by_p[i] >> 32) == p1 && (uint32_t)(by_p[j] >> 32) == p2) {
        const uint32_t a = (uint32_t)by_p[i], b = (uint32_t)by_p[j];
        if (a == b) return true;
        if (a < b) i++; else j++;
    }
    return false;
}
// one thread per pure subclonotype p0: partners = pures holding one of p0's CDR3s with at least mult * cells(p0) cells; p0 goes if two
// partners share no CDR3 with each other
__global__ void k_r_doublets(const u64 *by_c, const uint32_t *by_c_pure, const u64 *by_p, uint32_t m, const u64 *npure, uint32_t n_pure, int mult,
                             uint8_t *pdel, RCounters *ctr) {
    const size_t p0 = TID;
    if (p0 >= n_pure) return;
    const u64 need = (u64)mult * npure[p0];
    const uint32_t lo = lower_bound64(by_p, m, (u64)p0 << 32);
    u64 tested = 0;
    bool hit = false;
    for (uint32_t ia = lo; ia < m && (uint32_t)(by_p[ia] >> 32) == p0 && !hit; ia++) {
        const uint32_t ca = (uint32_t)by_p[ia];
        for (uint32_t ja = lower_bound64(by_c, m, (u64)ca << 32); ja < m && (uint32_t)(by_c[ja] >> 32) == ca && !hit; ja++) {
            const uint32_t p1 = by_c_pure[ja];
            if (npure[p1] < need) break;                  // sorted by cells, largest first
            if (p1 == p0) continue;
            for (uint32_t ib = ia + 1; ib < m && (uint32_t)(by_p[ib] >> 32) == p0 && !hit; ib++) {      // a second CDR3 of p0 (the same one would be shared by p1 and p2)
                const uint32_t cb = (uint32_t)by_p[ib];
                for (uint32_t jb = lower_bound64(by_c, m, (u64)cb << 32); jb < m && (uint32_t)(by_c[jb] >> 32) == cb; jb++) {
                    const uint32_t p2 = by_c_pure[jb];
                    if (npure[p2] < need) break;
                    if (p2 == p0 || p2 == p1) continue;
                    tested++;
                    if (!pures_share(by_p, m, p1, p2)) { hit = true; break; }
                }
            }
        }
    }
    if (hit) pdel[p0] = 1;
    if (tested) atomicAdd(&ctr->triples, tested);
}
__global__ void k_r_apply_pure_del(RDev d, const uint8_t *alive, const uint32_t *pure, const uint8_t *pdel, uint8_t *del, RCounters *ctr) {
    const size_t x = TID;
    if (x >= d.nx || !alive[x]) return;
    if (pdel[pure[x]]) { del[x] = 1; atomicAdd(&ctr->n_del, 1ull); }
}

// ---- F2: signature ----
__global__ void k_r_two_chain_cells(RDev d, const uint8_t *alive, const uint32_t *label, const uint32_t *col, const uint32_t *gbase, u64 *A) {
    const size_t x = TID;
    if (x >= d.nx || !alive[x]) return;
    const uint32_t o = d.chain_off[x];
    if (d.chain_off[x + 1] - o != 2) return;
    const uint32_t g = gbase[label[x]];
    atomicAdd(&A[g + col[o]], (u64)d.ncells[x]);
    atomicAdd(&A[g + col[o + 1]], (u64)d.ncells[x]);
}
__global__ void k_r_signature(RDev d, const uint8_t *alive, const uint32_t *label, const uint32_t *col, const uint32_t *gbase, const u64 *A, const uint32_t *pure,
                              const u64 *npure, const u64 *pure_hash, const RCounters *ctrc, int mult, uint8_t *del, RCounters *ctr) {
    const size_t x = TID;
    if (x >= d.nx || !alive[x]) return;
    const uint32_t o = d.chain_off[x], m = d.chain_off[x + 1] - o;
    if (m < 2) return;
    const uint32_t g = gbase[label[x]], np = (uint32_t)ctrc->n_pure;
    u64 tot = 0;
    for (uint32_t a = 0; a < m; a++) tot += A[g + col[o + a]];
    // a two-chain pure subclonotype inside p's columns was counted once per column: take one back (it is p itself when p has two chains)
    for (uint32_t a = 0; a < m; a++) for (uint32_t b = a + 1; b < m; b++) {
        const uint32_t c1 = min(col[o + a], col[o + b]), c2 = max(col[o + a], col[o + b]);
        u64 h = mix64(mix64(mix64(0x243F6A8885A308D3ull, label[x]), c1), c2);
        if (h == ~0ull) h = 0;
        const uint32_t q = lower_bound64(pure_hash, np, h);
        if (q < np && pure_hash[q] == h) tot -= npure[q];
    }
    if (m == 2) tot -= npure[pure[x]];
    if (tot >= (u64)mult * npure[pure[x]]) { del[x] = 1; atomicAdd(&ctr->n_del, 1ull); }
}

// ---- F3: weak chains ----
__global__ void k_r_column_cells(RDev d, const uint8_t *alive, const uint32_t *label, const uint32_t *col, const uint32_t *gbase, u64 *W, u64 *T) {
    const size_t x = TID;
    if (x >= d.nx || !alive[x]) return;
    const uint32_t o = d.chain_off[x], m = d.chain_off[x + 1] - o, g = gbase[label[x]];
    for (uint32_t a = 0; a < m; a++) atomicAdd(&W[g + col[o + a]], (u64)d.ncells[x]);
    atomicAdd(&T[label[x]], (u64)d.ncells[x]);
}
__global__ void k_r_weak(RDev d, const uint8_t *alive, const uint32_t *label, const uint32_t *col, const uint32_t *gbase, const uint32_t *ncols_label, const u64 *W,
                         const u64 *T, int max_cells, int mult, uint8_t *del, RCounters *ctr) {
    const size_t x = TID;
    if (x >= d.nx || !alive[x] || ncols_label[label[x]] < 3) return;
    const uint32_t o = d.chain_off[x], m = d.chain_off[x + 1] - o, g = gbase[label[x]];
    bool weak = false;
    for (uint32_t a = 0; a < m; a++) { const u64 w = W[g + col[o + a]]; weak |= w <= (u64)max_cells && (u64)mult * w < T[label[x]]; }
    if (weak) { del[x] = 1; atomicAdd(&ctr->n_del, 1ull); }
}
__global__ void k_r_commit(RDev d, const uint8_t *del, uint8_t code, uint8_t *alive, uint8_t *fate) {
    const size_t x = TID;
    if (x < d.nx && del[x]) { alive[x] = 0; fate[x] = code; }
}
__global__ void k_r_outputs(RDev d, const uint8_t *alive, const uint32_t *label, const uint32_t *ncols_label, uint32_t *clon, uint32_t *ncols, RCounters *ctr) {
    const size_t x = TID;
    if (x >= d.nx) return;
    clon[x] = alive[x] ? label[x] : NONE32;
    ncols[x] = alive[x] ? ncols_label[label[x]] : 0u;
    if (alive[x] && label[x] == x) atomicAdd(&ctr->n_del, 1ull);       // reused as the clonotype count
}

struct Pool {          // stream-ordered allocations, all returned at the end
    cudaStream_t st = nullptr; std::vector<void *> ptrs; bool failed = false; std::string msg; unsigned long long h2d = 0;
    template <typename T> T *get(size_t n) {
        void *p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T) + 64, st);
        if (e != cudaSuccess) { failed = true; msg = cudaGetErrorString(e); return nullptr; }
        ptrs.push_back(p);
        return (T *)p;
    }
    template <typename T> T *up(const T *h, size_t n) {
        T *p = get<T>(n);
        h2d += n * sizeof(T);
        if (p && n && cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) { failed = true; msg = "H2D copy"; }
        return p;
    }
    ~Pool() { if (!st) return; for (void *p : ptrs) cudaFreeAsync(p, st); cudaStreamSynchronize(st); cudaStreamDestroy(st); }
};

inline unsigned blocks(size_t n) { return (unsigned)std::max<size_t>(1, (n + 255) / 256); }

}  // namespace

extern "C" void erefine_opts_default(erefine_opts_t *o) {
    memset(o, 0, sizeof(*o));
    o->doublet = o->signature = o->weak_chains = 1;
    o->doublet_mult = 5; o->signature_mult = 20; o->weak_max_cells = 20; o->weak_mult = 8; o->third_chain_diffs = 10;
}
extern "C" const char *erefine_last_error(void) { return r_err.c_str(); }

extern "C" int erefine_run(const erefine_pack_t *pk, const erefine_opts_t *op, int device, erefine_result_t *out, erefine_stats_t *stats) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!pk || !op || !out) return rfail(-1, "null argument");
    if (pk->abi_version != EREFINE_ABI_VERSION) return rfail(-1, "abi_version mismatch");
    const uint32_t n = pk->n_info, nx = pk->n_exact, nc = pk->n_chain;
    const u64 nj = pk->n_joins;
    if ((n && (!pk->exact_id || !pk->exact_cols[0] || !pk->exact_cols[1] || !pk->class_of)) || (nj && (!pk->join_k1 || !pk->join_k2)) || !pk->chain_off ||
        (nx && (!pk->ncells || !pk->origin_rank || !out->fate || !out->clonotype_of || !out->n_columns)) ||
        (nc && (!pk->seq_id || !pk->cdr3_id || !pk->order_rank || !pk->left || !pk->seq_len || !pk->seq_off || !out->column_of)) || (pk->seq_arena_bytes && !pk->seq_arena))
        return rfail(-1, "a pack or result array the counts imply is NULL");
    if (nj >= 0xFFFFFFFFull) return rfail(-4, "more than 2^32 - 2 raw joins");
    if (pk->chain_off[nx] != nc) return rfail(-1, "chain_off[n_exact] != n_chain");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) { cudaGetLastError(); return rfail(-3, "no usable CUDA device (the refinement has no CPU fallback)"); }
    RCK(cudaSetDevice(device));
    erefine_stats_t S;
    memset(&S, 0, sizeof(S));
    if (nx == 0) { if (stats) *stats = S; return 0; }
    {
        cudaMemPool_t mp; unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    Pool P;
    RCK(cudaStreamCreateWithFlags(&P.st, cudaStreamNonBlocking));
    cudaStream_t st = P.st;
    cudaEvent_t ev0, ev1, evA, evB;
    RCK(cudaEventCreate(&ev0)); RCK(cudaEventCreate(&ev1)); RCK(cudaEventCreate(&evA)); RCK(cudaEventCreate(&evB));
    struct EvBox { cudaEvent_t a, b, c, d; ~EvBox() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); cudaEventDestroy(d); } } evbox{ ev0, ev1, evA, evB };
    bool host_gap = false;

    RDev d;
    d.n = n; d.nx = nx; d.nc = nc; d.nj = nj; d.seq_arena = nullptr; d.off2 = nullptr;
    d.exact_id = P.up(pk->exact_id, n); d.cols0 = P.up(pk->exact_cols[0], n); d.cols1 = P.up(pk->exact_cols[1], n); d.class_of = P.up(pk->class_of, n);
    d.jk1 = P.up(pk->join_k1, nj); d.jk2 = P.up(pk->join_k2, nj);
    d.chain_off = P.up(pk->chain_off, (size_t)nx + 1); d.ncells = P.up(pk->ncells, nx); d.origin_rank = P.up(pk->origin_rank, nx);
    d.seq_id = P.up(pk->seq_id, nc); d.cdr3_id = P.up(pk->cdr3_id, nc); d.order_rank = P.up(pk->order_rank, nc); d.left = P.up(pk->left, nc);
    d.seq_len = P.up(pk->seq_len, nc);          // the V..J bytes themselves: only the third chains', gathered below
    const size_t big = std::max<size_t>(std::max<size_t>(nc, nx), n) + 1;
    size_t cub_bytes = 0, tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (u64 *)nullptr, (u64 *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)big); cub_bytes = std::max(cub_bytes, tb);
    cub::DeviceRadixSort::SortKeys(nullptr, tb, (u64 *)nullptr, (u64 *)nullptr, (int)big); cub_bytes = std::max(cub_bytes, tb);
    cub::DeviceScan::InclusiveScan(nullptr, tb, (uint32_t *)nullptr, (uint32_t *)nullptr, RMax(), (int)big); cub_bytes = std::max(cub_bytes, tb);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)big); cub_bytes = std::max(cub_bytes, tb);
    cub::DeviceScan::InclusiveSum(nullptr, tb, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)big); cub_bytes = std::max(cub_bytes, tb);
    cub::DeviceSelect::Unique(nullptr, tb, (u64 *)nullptr, (u64 *)nullptr, (uint32_t *)nullptr, (int)big); cub_bytes = std::max(cub_bytes, tb);
    void *cub_tmp = P.get<char>(cub_bytes + 256);
    uint32_t *xof = P.get<uint32_t>(nc), *pos = P.get<uint32_t>(nc), *pos2 = P.get<uint32_t>(nc), *cp = P.get<uint32_t>(nc), *xp = P.get<uint32_t>(nx);
    uint32_t *label = P.get<uint32_t>(nx), *cmin = P.get<uint32_t>(n), *col = P.get<uint32_t>(nc), *base = P.get<uint32_t>(nc), *mm = P.get<uint32_t>(nc);
    uint32_t *gbase = P.get<uint32_t>(nx), *ncols_label = P.get<uint32_t>(nx), *pure = P.get<uint32_t>(nx);
    uint8_t *alive = P.get<uint8_t>(nx), *fate = P.get<uint8_t>(nx), *del = P.get<uint8_t>(nx), *pdel = P.get<uint8_t>(nx);
    u64 *k_in = P.get<u64>(big), *k_out = P.get<u64>(big), *k2_in = P.get<u64>(big), *k2_out = P.get<u64>(big), *rep = P.get<u64>(nc);
    uint32_t *v_in = P.get<uint32_t>(big), *v_out = P.get<uint32_t>(big), *t32a = P.get<uint32_t>(big), *t32b = P.get<uint32_t>(big), *t32c = P.get<uint32_t>(big);
    u64 *npure = P.get<u64>(nx), *pure_hash = P.get<u64>(nx), *A = P.get<u64>(nc), *T = P.get<u64>(nx), *by_c = P.get<u64>(big), *by_p = P.get<u64>(big);
    uint32_t *by_c_pure = P.get<uint32_t>(big), *n_uniq = P.get<uint32_t>(4);
    RCounters *ctr = P.get<RCounters>(1);
    uint32_t *o_clon = P.get<uint32_t>(nx), *o_ncols = P.get<uint32_t>(nx);
    uint32_t *seq_perm = P.get<uint32_t>(nc), *seq_head = P.get<uint32_t>(nc), *first = P.get<uint32_t>(nc), *off2 = P.get<uint32_t>((size_t)nc + 1);
    uint8_t *need8 = P.get<uint8_t>(nc), *third8 = P.get<uint8_t>(nj);
    if (P.failed) return rfail(-5, "device allocation: " + P.msg);
    u64 launches = 0;
    RCK(cudaEventRecord(ev0, st));
    RCK(cudaMemsetAsync(ctr, 0, sizeof(RCounters), st));
    RCK(cudaMemsetAsync(alive, 0, nx, st)); RCK(cudaMemsetAsync(fate, 0, nx, st));
    RCK(cudaMemsetAsync(label, 0xFF, 4 * (size_t)nx, st)); RCK(cudaMemsetAsync(cmin, 0xFF, 4 * (size_t)std::max<uint32_t>(n, 1), st));
    // set-up: chain -> exact, the two global chain orders, the clonotypes the join left
    k_r_xof<<<blocks(nx), 256, 0, st>>>(d, xof, ctr);
    k_r_check_joins<<<blocks(nj), 256, 0, st>>>(d, ctr);
    k_r_class_min<<<blocks(n), 256, 0, st>>>(d, cmin, ctr);
    k_r_label0<<<blocks(n), 256, 0, st>>>(d, cmin, label, alive);
    launches += 4;
    RCounters hc;
    RCK(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    if (hc.err & RERR_RANGE) return rfail(-1, "an index of the pack lies outside its table (exact_id, class_of, exact_cols, chain_off, join_k1 / join_k2, seq_off) or an exact subclonotype has more than EREFINE_MAX_CHAINS chains");
    if (nc) {
        k_r_key_origin<<<blocks(nc), 256, 0, st>>>(d, xof, k_in, v_in);
        tb = cub_bytes; cub::DeviceRadixSort::SortPairs(cub_tmp, tb, k_in, k_out, v_in, v_out, (int)nc, 0, 64, st);
        k_r_scatter_pos<<<blocks(nc), 256, 0, st>>>(v_out, nc, pos2);
        k_r_key_order<<<blocks(nc), 256, 0, st>>>(d, pos2, k_in, v_in);
        tb = cub_bytes; cub::DeviceRadixSort::SortPairs(cub_tmp, tb, k_in, k_out, v_in, v_out, (int)nc, 0, 64, st);
        k_r_scatter_pos<<<blocks(nc), 256, 0, st>>>(v_out, nc, pos);
        launches += 4;
    }
    int lb = 33;                                  // bits of a (clonotype name << 32 | 32-bit field) key; dead entries are all ones and sort last
    while (lb < 64 && (1ull << (lb - 32)) < (u64)nx + 1) lb++;
    int pb = 1;                                   // bits of a chain position
    while (pb < 32 && (1ull << pb) < (u64)nc + 1) pb++;
    const int ob = pb + (lb - 32);                // bits of an orbit key (clonotype name << pb | position)
    // S3 once: which chains' bytes are needed, those bytes gathered on the host in chain order, the verdict per join
    RCK(cudaMemsetAsync(third8, 0, std::max<u64>(nj, 1), st));
    static thread_local std::vector<uint8_t> need_h, gathered;      // kept between calls: no page faults on a second call's gather
    if (nj && nc) {
        RCK(cudaMemsetAsync(need8, 0, nc, st));
        k_r_third_need<<<blocks(nj), 256, 0, st>>>(d, need8);
        k_r_need_len<<<blocks(nc), 256, 0, st>>>(d, need8, t32a);
        tb = cub_bytes; cub::DeviceScan::ExclusiveSum(cub_tmp, tb, t32a, off2, (int)nc, st);
        if (need_h.size() < nc) need_h.resize(nc);
        RCK(cudaMemcpyAsync(need_h.data(), need8, nc, cudaMemcpyDeviceToHost, st));
        RCK(cudaEventRecord(evA, st));
        RCK(cudaStreamSynchronize(st));
        host_gap = true;
        u64 total = 0;
        for (uint32_t c = 0; c < nc; c++) if (need_h[c]) {
            if (pk->seq_off[c] + pk->seq_len[c] > pk->seq_arena_bytes) return rfail(-1, "seq_off + seq_len of a chain lies outside seq_arena");
            total += pk->seq_len[c];
        }
        if (total >= 0xFFFFFFFFull) return rfail(-4, "third chains of joined three-chain exact subclonotypes hold 4 GiB of sequence or more");
        if (gathered.size() < total + 16) gathered.resize(total + 16);
        u64 o = 0;
        for (uint32_t c = 0; c < nc; c++) if (need_h[c]) { memcpy(gathered.data() + o, pk->seq_arena + pk->seq_off[c], pk->seq_len[c]); o += pk->seq_len[c]; }
        RCK(cudaEventRecord(evB, st));          // the host's gather is not device time
        d.seq_arena = P.up(gathered.data(), (size_t)total + 16);
        d.off2 = off2;
        if (P.failed) return rfail(-5, "device allocation: " + P.msg);
        if (total) k_r_third<<<blocks(nj), 256, 0, st>>>(d, op->third_chain_diffs, third8);
        launches += 3;
    }
    // S4 once: chains sorted by (clonotype the join left, sequence id), run heads
    if (nc) {
        k_r_seq_keys<<<blocks(nc), 256, 0, st>>>(d, xof, alive, label, k_in, v_in);
        tb = cub_bytes; cub::DeviceRadixSort::SortPairs(cub_tmp, tb, k_in, k_out, v_in, seq_perm, (int)nc, 0, lb, st);
        k_r_heads<<<blocks(nc), 256, 0, st>>>(k_out, nc, ~0ull, t32a);
        tb = cub_bytes; cub::DeviceScan::InclusiveScan(cub_tmp, tb, t32a, seq_head, RMax(), (int)nc, st);
        launches += 2;
    }
    auto stage = [&]() -> int {
        RCK(cudaMemsetAsync(&ctr->n_pure, 0, 8, st));
        k_r_iota<<<blocks(nc), 256, 0, st>>>(cp, nc);
        k_r_iota<<<blocks(nx), 256, 0, st>>>(xp, nx);
        if (nj) k_r_joins<<<blocks(nj), 256, 0, st>>>(d, alive, third8, cp);
        // S4
        RCK(cudaMemsetAsync(first, 0xFF, 4 * (size_t)nc, st));
        k_r_onesie_first<<<blocks(nc), 256, 0, st>>>(d, seq_perm, xof, alive, seq_head, first);
        k_r_onesie_join<<<blocks(nc), 256, 0, st>>>(d, seq_perm, xof, alive, seq_head, first, cp);
        // S6
        k_r_link_exacts<<<blocks(nc), 256, 0, st>>>(d, xof, alive, cp, xp);
        k_r_labels<<<blocks(nx), 256, 0, st>>>(d, alive, xp, label);
        // S5
        RCK(cudaMemsetAsync(rep, 0xFF, 8 * (size_t)nc, st)); RCK(cudaMemsetAsync(mm, 0, 4 * (size_t)nc, st));
        RCK(cudaMemsetAsync(ncols_label, 0, 4 * (size_t)nx, st)); RCK(cudaMemsetAsync(gbase, 0, 4 * (size_t)nx, st));
        k_r_orbit_stats<<<blocks(nx), 256, 0, st>>>(d, alive, cp, pos, pos2, rep, mm);
        k_r_orbit_keys<<<blocks(nc), 256, 0, st>>>(d, xof, alive, label, cp, rep, pb, k_in, v_in);
        tb = cub_bytes; cub::DeviceRadixSort::SortPairs(cub_tmp, tb, k_in, k_out, v_in, v_out, (int)nc, 0, ob, st);
        k_r_gather_mm<<<blocks(nc), 256, 0, st>>>(d, k_out, v_out, mm, t32c);
        tb = cub_bytes; cub::DeviceScan::ExclusiveSum(cub_tmp, tb, t32c, t32a, (int)nc, st);
        k_r_heads<<<blocks(nc), 256, 0, st>>>(k_out, nc, ~((1ull << pb) - 1ull), v_in);
        tb = cub_bytes; cub::DeviceScan::InclusiveScan(cub_tmp, tb, v_in, t32b, RMax(), (int)nc, st);
        k_r_orbit_base<<<blocks(nc), 256, 0, st>>>(d, k_out, v_out, t32b, t32a, t32c, pb, base, gbase, ncols_label);
        // columns, S7
        k_r_columns<<<blocks(nx), 256, 0, st>>>(d, alive, label, cp, base, col, k2_in, v_in);
        tb = cub_bytes; cub::DeviceRadixSort::SortPairs(cub_tmp, tb, k2_in, k2_out, v_in, v_out, (int)nx, 0, 64, st);
        k_r_pure_heads<<<blocks(nx), 256, 0, st>>>(d, k2_out, v_out, label, col, t32a, ctr);
        tb = cub_bytes; cub::DeviceScan::InclusiveSum(cub_tmp, tb, t32a, t32b, (int)nx, st);
        RCK(cudaMemsetAsync(npure, 0, 8 * (size_t)nx, st));
        k_r_pure_ids<<<blocks(nx), 256, 0, st>>>(d, k2_out, v_out, t32b, pure, npure, pure_hash, ctr);
        launches += 15 + (nj ? 1 : 0);
        RCK(cudaGetLastError());
        return 0;
    };
    auto finish_filter = [&](int code, uint64_t *count) -> int {      // counter to the host, deletions committed, the next stage
        RCK(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
        RCK(cudaStreamSynchronize(st));
        if (hc.err & RERR_HASH) return rfail(-4, "two column signatures collide in the 64-bit hash that groups pure subclonotypes");
        *count = hc.n_del;
        if (hc.n_del) {
            k_r_commit<<<blocks(nx), 256, 0, st>>>(d, del, (uint8_t)code, alive, fate);
            launches += 1;
            RCK(cudaMemsetAsync(&ctr->n_del, 0, 8, st));
            return stage();
        }
        return 0;
    };
    {   // clonotypes coming in (the classes of class_of, before any break-up): names that are their own label
        RCK(cudaMemsetAsync(ncols_label, 0, 4 * (size_t)nx, st));
        k_r_outputs<<<blocks(nx), 256, 0, st>>>(d, alive, label, ncols_label, o_clon, o_ncols, ctr);
        RCK(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
        RCK(cudaStreamSynchronize(st));
        S.n_clonotypes_in = hc.n_del;
        RCK(cudaMemsetAsync(&ctr->n_del, 0, 8, st));
        launches += 1;
    }
    if (int rc = stage()) return rc;
    RCK(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    S.n_pure = hc.n_pure;
    if (op->doublet) {
        RCK(cudaMemsetAsync(del, 0, nx, st)); RCK(cudaMemsetAsync(pdel, 0, nx, st));
        k_r_content<<<blocks(nc), 256, 0, st>>>(d, xof, alive, pure, k_in);
        tb = cub_bytes; cub::DeviceRadixSort::SortKeys(cub_tmp, tb, k_in, k_out, (int)nc, 0, 64, st);
        tb = cub_bytes; cub::DeviceSelect::Unique(cub_tmp, tb, k_out, k_in, n_uniq, (int)nc, st);
        k_r_content_by_cells<<<blocks(nc), 256, 0, st>>>(k_in, n_uniq, npure, k2_in, v_in, k_out, nc);
        tb = cub_bytes; cub::DeviceRadixSort::SortPairs(cub_tmp, tb, k2_in, by_c, v_in, by_c_pure, (int)nc, 0, 64, st);
        tb = cub_bytes; cub::DeviceRadixSort::SortKeys(cub_tmp, tb, k_out, by_p, (int)nc, 0, lb, st);
        // the lists hold n_uniq live entries followed by ~0 keys (which sort last); the kernels search the whole arrays
        uint32_t np_host = (uint32_t)S.n_pure;
        k_r_doublets<<<blocks(np_host), 256, 0, st>>>(by_c, by_c_pure, by_p, nc, npure, np_host, op->doublet_mult, pdel, ctr);
        k_r_apply_pure_del<<<blocks(nx), 256, 0, st>>>(d, alive, pure, pdel, del, ctr);
        launches += 4;
        if (int rc = finish_filter(EREFINE_DOUBLET, &S.n_doublet)) return rc;
        S.triples_tested = hc.triples;
    }
    if (op->signature) {
        RCK(cudaMemsetAsync(del, 0, nx, st)); RCK(cudaMemsetAsync(A, 0, 8 * (size_t)nc, st));
        k_r_two_chain_cells<<<blocks(nx), 256, 0, st>>>(d, alive, label, col, gbase, A);
        k_r_signature<<<blocks(nx), 256, 0, st>>>(d, alive, label, col, gbase, A, pure, npure, pure_hash, ctr, op->signature_mult, del, ctr);
        launches += 2;
        if (int rc = finish_filter(EREFINE_SIGNATURE, &S.n_signature)) return rc;
    }
    if (op->weak_chains) {
        RCK(cudaMemsetAsync(del, 0, nx, st)); RCK(cudaMemsetAsync(A, 0, 8 * (size_t)nc, st)); RCK(cudaMemsetAsync(T, 0, 8 * (size_t)nx, st));
        k_r_column_cells<<<blocks(nx), 256, 0, st>>>(d, alive, label, col, gbase, A, T);
        k_r_weak<<<blocks(nx), 256, 0, st>>>(d, alive, label, col, gbase, ncols_label, A, T, op->weak_max_cells, op->weak_mult, del, ctr);
        launches += 2;
        if (int rc = finish_filter(EREFINE_WEAK_CHAIN, &S.n_weak)) return rc;
    }
    RCK(cudaMemsetAsync(&ctr->n_del, 0, 8, st));
    k_r_outputs<<<blocks(nx), 256, 0, st>>>(d, alive, label, ncols_label, o_clon, o_ncols, ctr);
    launches += 1;
    RCK(cudaEventRecord(ev1, st));
    RCK(cudaMemcpyAsync(out->fate, fate, nx, cudaMemcpyDeviceToHost, st));
    RCK(cudaMemcpyAsync(out->clonotype_of, o_clon, 4 * (size_t)nx, cudaMemcpyDeviceToHost, st));
    RCK(cudaMemcpyAsync(out->n_columns, o_ncols, 4 * (size_t)nx, cudaMemcpyDeviceToHost, st));
    if (nc) RCK(cudaMemcpyAsync(out->column_of, col, 4 * (size_t)nc, cudaMemcpyDeviceToHost, st));
    RCK(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    RCK(cudaGetLastError());
    S.n_clonotypes_out = hc.n_del;
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    if (host_gap) { float gap = 0; cudaEventElapsedTime(&gap, evA, evB); ms -= gap; }
    S.ms_device = ms;
    S.h2d_bytes = P.h2d;
    S.gpu_launches = launches;
    S.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = S;
    return 0;
}
