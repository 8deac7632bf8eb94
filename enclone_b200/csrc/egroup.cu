// egroup.cu — symmetric clonotype grouping on the GPU (include/enclone_group.h).
//
// Replaces the distance passes of enclone_tail::grouper::grouper(), Case 1
// (/root/reference/enclone_tail/src/grouper.rs:386-758): inside every group, all pairs of clonotypes are tested
// with a sequence-distance criterion and the accepted pairs are merged in an EquivRel.  The reference walks the pair
// triangle sequentially, skipping pairs already in one class; the classes are the connected components of the accept
// graph whatever the order, so here every pair is a thread: bit-parallel Levenshtein (Myers / Hyyro block algorithm,
// 64 rows of the DP matrix per word, the vertical deltas Pv / Mv in registers or local memory), GPU union-find with
// compare-and-swap hooking.  Key partitions, the orbit bookkeeping between passes (with the reference's own
// `ee.orbit(i)` idiom, grouper.rs:441-452) and the final joins (:760-766) are host code in this file.
// No CPU fallback for the distance passes: without a device egroup_run fails with -3.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/enclone_group.h"
#include "egroup_myers.h"

namespace {

thread_local std::string g_err;
int gfail(int code, const std::string &m) { g_err = m; return code; }
#define GCK(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return gfail(-2, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

enum { D_SEQ = 0, D_HF = 1, D_AA = 2, D_CDR3 = 3, D_CDR3AA = 4 };

struct GDev {                      // device copies of the pack's tables
    const uint32_t *clono_exact_off, *clono_exact, *exact_chain_off;
    const uint8_t *left;
    const uint64_t *seq_off; const uint32_t *seq_len, *cdr3_start;
    const uint64_t *cdr3_aa_off; const uint32_t *cdr3_aa_len;
    const uint8_t *seq_arena, *aa_arena;
};
struct GCounters { unsigned long long pairs_evaluated, chain_compares; unsigned int err; };

#define ERR_TOO_LONG 1u
#define ERR_BAD_BYTE 2u

__device__ __forceinline__ uint32_t uf_find(uint32_t *p, uint32_t x) {
    for (;;) {
        const uint32_t px = __ldcg(&p[x]);
        if (px == x) return x;
        const uint32_t gp = __ldcg(&p[px]);
        if (gp != px) p[x] = gp;
        x = px;
    }
}
__device__ __forceinline__ void uf_union(uint32_t *p, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find(p, a); b = uf_find(p, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        if (atomicCAS(&p[a], a, b) == a) return;
    }
}

// amino::nucleotide_to_aminoacid_sequence(dna, 0) (grouper.rs:576-580)
__device__ uint32_t translate(const uint8_t *dna, uint32_t n, uint8_t *aa) {
    const char *tab = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";
    uint32_t k = 0;
    for (uint32_t i = 0; i + 3 <= n; i += 3) {
        uint32_t idx = 0;
        bool ok = true;
        for (int q = 0; q < 3; q++) { const uint32_t c = sym_nt(dna[i + q]); ok = ok && c < 4; idx = idx * 4 + (c & 3); }
        aa[k++] = ok ? (uint8_t)tab[idx] : (uint8_t)'X';
    }
    return k;
}

// one chain pair against the criterion of pass `kind` (grouper.rs:426-436, 499-512, 576-590, 648-664, 724-735)
__device__ bool chain_pair_ok(const GDev &d, int kind, double min_r, const double *__restrict__ penalty, uint32_t h1, uint32_t h2, GCounters *ctr) {
    const uint8_t *s1 = d.seq_arena + d.seq_off[h1], *s2 = d.seq_arena + d.seq_off[h2];
    const uint32_t n1 = d.seq_len[h1], n2 = d.seq_len[h2];
    const uint8_t *a1 = d.aa_arena + d.cdr3_aa_off[h1], *a2 = d.aa_arena + d.cdr3_aa_off[h2];
    const uint32_t m1 = d.cdr3_aa_len[h1], m2 = d.cdr3_aa_len[h2];
    switch (kind) {
        case D_SEQ: {
            if (min(n1, n2) > 64 * MAXW) { atomicOr(&ctr->err, ERR_TOO_LONG); return false; }
            const int dm = dmax_ratio(n1, n2, min_r);               // r1 >= min_r || r2 >= min_r  <=>  distance <= dm
            return dm >= 0 && lev_within<7, false>(s1, n1, s2, n2, (uint32_t)dm);
        }
        case D_AA: {
            if (min(n1, n2) / 3 > 64 * MAXW || max(n1, n2) > 3 * 512) { atomicOr(&ctr->err, ERR_TOO_LONG); return false; }
            uint8_t t1[512], t2[512];
            const uint32_t k1 = translate(s1, n1, t1), k2 = translate(s2, n2, t2);
            const int dm = dmax_ratio(k1, k2, min_r);
            return dm >= 0 && lev_within<32, true>(t1, k1, t2, k2, (uint32_t)dm);
        }
        case D_HF: {
            if (m1 != m2) return false;
            double err = 0.0;
            for (uint32_t k = 0; k < m1; k++) {
                const unsigned x = (unsigned)(a1[k] - 'A'), y = (unsigned)(a2[k] - 'A');
                err = __dadd_rn(err, penalty[(x < 27 ? x : 26) * 27 + (y < 27 ? y : 26)]);
            }
            err = __ddiv_rn(err, (double)m1);
            return err <= __dsub_rn(1.0, min_r);
        }
        case D_CDR3: {
            const uint32_t l1 = 3 * m1, l2 = 3 * m2;
            if (min(l1, l2) > 64 * MAXW) { atomicOr(&ctr->err, ERR_TOO_LONG); return false; }
            const double d_max_f = __dmul_rn(__dsub_rn(1.0, min_r), (double)min(l1, l2));
            const uint32_t d_max = (uint32_t)floor(d_max_f);
            return lev_within<7, false>(s1 + d.cdr3_start[h1], l1, s2 + d.cdr3_start[h2], l2, d_max);
        }
        default: {
            if (min(m1, m2) > 64 * MAXW) { atomicOr(&ctr->err, ERR_TOO_LONG); return false; }
            const double d_max_f = __dmul_rn(__dsub_rn(1.0, min_r), (double)min(m1, m2));
            const uint32_t d_max = (uint32_t)floor(d_max_f);
            return lev_within<32, true>(a1, m1, a2, m2, d_max);
        }
    }
}

// One thread per clonotype pair (i1 < i2) of a group; `members` = all groups concatenated, grp_off / pair_off their
// member and pair prefix sums.  A pair already in one class is skipped (as the reference does; it cannot change the
// classes), the others run the loop nest of grouper.rs:406-438.
__global__ void __launch_bounds__(128) k_group_pairs(GDev d, int kind, int want_left, double min_r, const double *__restrict__ penalty,
                                                     const uint32_t *__restrict__ members, const uint32_t *__restrict__ grp_off,
                                                     const unsigned long long *__restrict__ pair_off, uint32_t n_groups, unsigned long long n_pairs,
                                                     uint32_t *label, GCounters *ctr) {
    unsigned long long evaluated = 0, compares = 0;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n_groups;                              // group of pair t
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (pair_off[mid] <= t) lo = mid; else hi = mid; }
        const unsigned long long r = t - pair_off[lo];
        // pair r of the triangle, row-major over i2: r = i2 (i2 - 1) / 2 + i1, i1 < i2
        uint32_t i2 = (uint32_t)((1.0 + sqrt(1.0 + 8.0 * (double)r)) * 0.5);
        while ((unsigned long long)i2 * (i2 - 1) / 2 > r) i2--;
        while ((unsigned long long)(i2 + 1) * i2 / 2 <= r) i2++;
        const uint32_t i1 = (uint32_t)(r - (unsigned long long)i2 * (i2 - 1) / 2);
        const uint32_t p1 = grp_off[lo] + i1, p2 = grp_off[lo] + i2;
        if (uf_find(label, p1) == uf_find(label, p2)) continue;
        evaluated++;
        const uint32_t g1 = members[p1], g2 = members[p2];
        bool ok = false;
        for (uint32_t x1 = d.clono_exact_off[g1]; x1 < d.clono_exact_off[g1 + 1] && !ok; x1++)
            for (uint32_t x2 = d.clono_exact_off[g2]; x2 < d.clono_exact_off[g2 + 1] && !ok; x2++) {
                const uint32_t u1 = d.clono_exact[x1], u2 = d.clono_exact[x2];
                for (uint32_t c1 = d.exact_chain_off[u1]; c1 < d.exact_chain_off[u1 + 1] && !ok; c1++) {
                    if ((want_left != 0) != (d.left[c1] != 0)) continue;
                    for (uint32_t c2 = d.exact_chain_off[u2]; c2 < d.exact_chain_off[u2 + 1] && !ok; c2++) {
                        if ((want_left != 0) != (d.left[c2] != 0)) continue;
                        compares++;
                        ok = chain_pair_ok(d, kind, min_r, penalty, c1, c2, ctr);
                    }
                }
            }
        if (ok) uf_union(label, p1, p2);
    }
    for (int o = 16; o > 0; o >>= 1) { evaluated += __shfl_xor_sync(0xFFFFFFFFu, evaluated, o); compares += __shfl_xor_sync(0xFFFFFFFFu, compares, o); }
    if ((threadIdx.x & 31) == 0 && (evaluated | compares)) { atomicAdd(&ctr->pairs_evaluated, evaluated); atomicAdd(&ctr->chain_compares, compares); }
}
// every byte the nucleotide criteria compare must be one of ACGT, N, - (the symbol map of the bit-parallel kernel); 0 pads
__global__ void k_validate_nt(const uint8_t *__restrict__ a, unsigned long long n, GCounters *ctr) {
    bool bad = false;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) bad = bad || (a[i] != 0 && sym_nt(a[i]) == 6);
    if (bad) atomicOr(&ctr->err, ERR_BAD_BYTE);
}
__global__ void k_iota(uint32_t *p, uint32_t n) { const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }
__global__ void k_flatten(uint32_t *label, uint32_t n, uint32_t *out) { const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = uf_find(label, i); }


// ---- Case 2: asymmetric grouping (grouper.rs:828-928) ---------------------------------------------------------------
#define ASYM_INF 1000000000u
#define ASYM_BINS 2048             // distances of real pairs are <= 2 * 512; everything else is ASYM_INF

// distance(center c, clonotype j), grouper.rs:856-889: integers throughout (the source's f64 values are these integers)
__device__ uint32_t asym_dist(const GDev &d, uint32_t c, uint32_t j, GCounters *ctr, unsigned long long &compares) {
    uint32_t best = ASYM_INF;
    for (uint32_t x1 = d.clono_exact_off[c]; x1 < d.clono_exact_off[c + 1]; x1++)
        for (uint32_t x2 = d.clono_exact_off[j]; x2 < d.clono_exact_off[j + 1]; x2++) {
            const uint32_t u1 = d.clono_exact[x1], u2 = d.clono_exact[x2];
            uint32_t heavy = ASYM_INF, light = ASYM_INF;
            for (uint32_t m1 = d.exact_chain_off[u1]; m1 < d.exact_chain_off[u1 + 1]; m1++)
                for (uint32_t m2 = d.exact_chain_off[u2]; m2 < d.exact_chain_off[u2 + 1]; m2++) {
                    if ((d.left[m1] != 0) != (d.left[m2] != 0)) continue;
                    const uint32_t l1 = d.cdr3_aa_len[m1], l2 = d.cdr3_aa_len[m2];
                    if (min(l1, l2) > 64 * MAXW || max(l1, l2) > 512) { atomicOr(&ctr->err, ERR_TOO_LONG); continue; }
                    const uint32_t x = lev_aa(d.aa_arena + d.cdr3_aa_off[m1], l1, d.aa_arena + d.cdr3_aa_off[m2], l2);
                    compares++;
                    if (d.left[m1]) heavy = min(heavy, x); else light = min(light, x);
                }
            best = min(best, heavy + light);
        }
    return best;
}

// The center's side of every comparison is the same for a whole row of the distance matrix: its chains' match vectors are
// built once per CTA into shared memory (CDR3s of up to 64 amino acids, up to ASYM_MAXCH chains; anything else takes the
// generic path), and a comparison is then the one-word recurrence over the other clonotype's string alone.
#define ASYM_MAXCH 16
struct CenterCache {
    unsigned long long peq[ASYM_MAXCH][32];
    uint32_t present[ASYM_MAXCH], gid[ASYM_MAXCH];
    uint16_t len[ASYM_MAXCH];
    uint8_t left[ASYM_MAXCH], ex_off[ASYM_MAXCH + 1];
    uint32_t n_ex, n_ch;
    int fast;
};
__device__ void load_center(const GDev &d, uint32_t c, CenterCache &cc) {        // block-wide; ends with a barrier
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t nch = 0, nex = 0;
        bool ok = true;
        for (uint32_t x = d.clono_exact_off[c]; x < d.clono_exact_off[c + 1] && ok; x++) {
            const uint32_t u = d.clono_exact[x];
            if (nex >= ASYM_MAXCH) { ok = false; break; }
            cc.ex_off[nex++] = (uint8_t)nch;
            for (uint32_t m = d.exact_chain_off[u]; m < d.exact_chain_off[u + 1]; m++) {
                if (nch >= ASYM_MAXCH || d.cdr3_aa_len[m] > 64) { ok = false; break; }
                cc.gid[nch] = m; cc.len[nch] = (uint16_t)d.cdr3_aa_len[m]; cc.left[nch] = d.left[m] != 0; nch++;
            }
        }
        cc.ex_off[nex < ASYM_MAXCH ? nex : ASYM_MAXCH] = (uint8_t)nch;
        cc.n_ex = nex; cc.n_ch = nch; cc.fast = ok ? 1 : 0;
    }
    __syncthreads();
    if (cc.fast && threadIdx.x < cc.n_ch) {
        const uint32_t t = threadIdx.x, m = cc.gid[t];
        const uint8_t *p = d.aa_arena + d.cdr3_aa_off[m];
        uint32_t present = 0;
        for (uint32_t i = 0; i < cc.len[t]; i++) {
            const uint32_t sy = sym_aa(p[i]);
            if (!((present >> sy) & 1u)) { cc.peq[t][sy] = 0ull; present |= 1u << sy; }
            cc.peq[t][sy] |= 1ull << i;
        }
        cc.present[t] = present;
    }
    __syncthreads();
}
__device__ uint32_t asym_dist_cached(const GDev &d, const CenterCache &cc, uint32_t c, uint32_t j, GCounters *ctr, unsigned long long &compares) {
    if (!cc.fast) return asym_dist(d, c, j, ctr, compares);
    uint32_t best = ASYM_INF;
    for (uint32_t e = 0; e < cc.n_ex; e++)
        for (uint32_t x2 = d.clono_exact_off[j]; x2 < d.clono_exact_off[j + 1]; x2++) {
            const uint32_t u2 = d.clono_exact[x2];
            uint32_t heavy = ASYM_INF, light = ASYM_INF;
            for (uint32_t ci = cc.ex_off[e]; ci < cc.ex_off[e + 1]; ci++)
                for (uint32_t m2 = d.exact_chain_off[u2]; m2 < d.exact_chain_off[u2 + 1]; m2++) {
                    if ((cc.left[ci] != 0) != (d.left[m2] != 0)) continue;
                    const uint32_t l2 = d.cdr3_aa_len[m2];
                    if (l2 > 512) { atomicOr(&ctr->err, ERR_TOO_LONG); continue; }
                    const uint32_t x = myers64_aa(cc.peq[ci], cc.present[ci], cc.len[ci], d.aa_arena + d.cdr3_aa_off[m2], l2);
                    compares++;
                    if (cc.left[ci]) heavy = min(heavy, x); else light = min(light, x);
                }
            best = min(best, heavy + light);
        }
    return best;
}

// max=D / no bound: a CTA takes (center, slice of the clonotypes); the pairs within the bound are appended to one list
__global__ void __launch_bounds__(256) k_asym_max(GDev d, const uint32_t *__restrict__ center, uint32_t n_center, uint32_t n, uint32_t slices, uint32_t dcut,
                                                  uint4 *out, unsigned long long cap, unsigned long long *n_out, GCounters *ctr) {
    __shared__ CenterCache cc;
    unsigned long long compares = 0, pairs = 0;
    const uint32_t per = (n + slices - 1) / slices;
    uint32_t loaded = 0xFFFFFFFFu;
    for (unsigned long long vb = blockIdx.x; vb < (unsigned long long)n_center * slices; vb += gridDim.x) {
        const uint32_t i = (uint32_t)(vb / slices), sl = (uint32_t)(vb % slices);
        if (i != loaded) { load_center(d, center[i], cc); loaded = i; }
        const uint32_t j1 = min(n, (sl + 1) * per);
        for (uint32_t j = sl * per + threadIdx.x; j < j1; j += blockDim.x) {
            const uint32_t dist = asym_dist_cached(d, cc, center[i], j, ctr, compares);
            pairs++;
            if (dist <= dcut) {
                const unsigned long long w = atomicAdd(n_out, 1ull);
                if (w < cap) out[w] = make_uint4(i, j, dist, 0u);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { compares += __shfl_xor_sync(0xFFFFFFFFu, compares, o); pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, o); }
    if ((threadIdx.x & 31) == 0 && (compares | pairs)) { atomicAdd(&ctr->chain_compares, compares); atomicAdd(&ctr->pairs_evaluated, pairs); }
}

// top=N: one CTA per center.  Pass 1 computes the center's row of distances (kept in `row`) and a histogram of them in
// shared memory; the histogram gives the distance d* at which the sorted order reaches position N and how many entries of
// distance d* are still needed; pass 2 walks the row in ascending index order and emits everything below d* plus the first
// `r` entries at d* — the N+1 smallest by (distance, index), the order of id.sort_by at grouper.rs:908.
__global__ void __launch_bounds__(256) k_asym_top(GDev d, const uint32_t *__restrict__ center, uint32_t n_center, uint32_t n, uint32_t need,
                                                  uint32_t *row_all, uint2 *out, GCounters *ctr) {
    __shared__ uint32_t hist[ASYM_BINS];
    __shared__ uint32_t s_dstar, s_r, s_base_take, s_base_eq;
    __shared__ uint32_t scan_take[256], scan_eq[256];
    __shared__ CenterCache cc;
    uint32_t *row = row_all + (size_t)blockIdx.x * n;
    unsigned long long compares = 0, pairs = 0;
    for (uint32_t i = blockIdx.x; i < n_center; i += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < ASYM_BINS; b += blockDim.x) hist[b] = 0;
        const uint32_t c = center[i];
        load_center(d, c, cc);
        for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
            const uint32_t dist = asym_dist_cached(d, cc, c, j, ctr, compares);
            pairs++;
            row[j] = dist;
            atomicAdd(&hist[min(dist, (uint32_t)ASYM_BINS - 1)], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t cum = 0, b = 0;
            for (; b < ASYM_BINS; b++) { if (cum + hist[b] >= need) break; cum += hist[b]; }
            s_dstar = b; s_r = need - cum; s_base_take = 0; s_base_eq = 0;
        }
        __syncthreads();
        const uint32_t dstar = s_dstar, r = s_r;
        for (uint32_t j0 = 0; j0 < n; j0 += blockDim.x) {
            const uint32_t j = j0 + threadIdx.x;
            const uint32_t code = j < n ? min(row[j], (uint32_t)ASYM_BINS - 1) : 0xFFFFFFFFu;
            const uint32_t eq = code == dstar ? 1u : 0u, below = code < dstar ? 1u : 0u;
            scan_eq[threadIdx.x] = eq;
            __syncthreads();
            // inclusive scans over the chunk (Hillis-Steele; 256 entries)
            for (uint32_t o = 1; o < blockDim.x; o <<= 1) { const uint32_t v = threadIdx.x >= o ? scan_eq[threadIdx.x - o] : 0u; __syncthreads(); scan_eq[threadIdx.x] += v; __syncthreads(); }
            const uint32_t eq_rank = s_base_eq + scan_eq[threadIdx.x] - eq;           // entries at d* before this one
            const uint32_t take = below | (eq && eq_rank < r ? 1u : 0u);
            scan_take[threadIdx.x] = take;
            __syncthreads();
            for (uint32_t o = 1; o < blockDim.x; o <<= 1) { const uint32_t v = threadIdx.x >= o ? scan_take[threadIdx.x - o] : 0u; __syncthreads(); scan_take[threadIdx.x] += v; __syncthreads(); }
            if (take) out[(size_t)i * need + s_base_take + scan_take[threadIdx.x] - 1] = make_uint2(j, row[j]);
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1) { s_base_take += scan_take[threadIdx.x]; s_base_eq += scan_eq[threadIdx.x]; }
            __syncthreads();
        }
    }
    for (int o = 16; o > 0; o >>= 1) { compares += __shfl_xor_sync(0xFFFFFFFFu, compares, o); pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, o); }
    if ((threadIdx.x & 31) == 0 && (compares | pairs)) { atomicAdd(&ctr->chain_compares, compares); atomicAdd(&ctr->pairs_evaluated, pairs); }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef std::vector<std::vector<uint32_t>> Groups;

enum { K_VJ, K_VDJ, K_VH, K_VJH, K_VDJH, K_VJLEN, K_CDR3LEN, K_CDR3HLEN, K_CDR3LLEN };

// the key of clonotype i for a key partition (grouper.rs:64-384), appended to `k`; returns `lonely` = no heavy chain in a
// *_heavy_refname pass.  Names are interned ids: equal strings <=> equal ids, which is all the partition depends on.
bool append_key(const egroup_pack_t *pk, int kind, uint32_t i, std::vector<int64_t> &k) {
    const size_t at = k.size();
    const uint32_t c0 = pk->col_off[i], c1 = pk->col_off[i + 1];
    const uint32_t ex0 = pk->clono_exact[pk->clono_exact_off[i]];
    const uint32_t h0 = pk->exact_chain_off[ex0], h1 = pk->exact_chain_off[ex0 + 1];
    bool lonely = false;
    switch (kind) {
        case K_VJ: for (uint32_t c = c0; c < c1; c++) { k.push_back(pk->col_vname[c]); k.push_back(pk->col_jname[c]); } break;
        case K_VDJ: for (uint32_t c = c0; c < c1; c++) { k.push_back(pk->col_vname[c]); if (pk->col_left[c]) k.push_back(pk->col_dname[c]); k.push_back(pk->col_jname[c]); } break;
        case K_VH: for (uint32_t c = c0; c < c1; c++) if (pk->col_left[c]) k.push_back(pk->col_vname[c]); lonely = k.size() == at; break;
        case K_VJH: for (uint32_t c = c0; c < c1; c++) if (pk->col_left[c]) { k.push_back(pk->col_vname[c]); k.push_back(pk->col_jname[c]); } lonely = k.size() == at; break;
        case K_VDJH: for (uint32_t c = c0; c < c1; c++) if (pk->col_left[c]) { k.push_back(pk->col_vname[c]); k.push_back(pk->col_dname[c]); k.push_back(pk->col_jname[c]); } lonely = k.size() == at; break;
        case K_VJLEN: for (uint32_t h = h0; h < h1; h++) k.push_back(((int64_t)pk->left[h] << 32) | pk->seq_del_len[h]); std::sort(k.begin() + at, k.end()); break;
        case K_CDR3LEN: for (uint32_t h = h0; h < h1; h++) k.push_back(((int64_t)pk->left[h] << 32) | pk->cdr3_aa_len[h]); std::sort(k.begin() + at, k.end()); break;
        case K_CDR3HLEN: for (uint32_t h = h0; h < h1; h++) if (pk->left[h]) k.push_back(pk->cdr3_aa_len[h]); std::sort(k.begin() + at, k.end()); break;
        case K_CDR3LLEN: for (uint32_t h = h0; h < h1; h++) if (!pk->left[h]) k.push_back(pk->cdr3_aa_len[h]); std::sort(k.begin() + at, k.end()); break;
    }
    return lonely;
}

// one key partition: inside every group sort by (key, index), runs of equal keys become the new groups
void key_pass(const egroup_pack_t *pk, int kind, Groups &G) {
    Groups out;
    std::vector<int64_t> buf;
    std::vector<uint32_t> off, order;
    std::vector<uint8_t> lonely;
    std::vector<std::pair<uint64_t, uint64_t>> hk;
    for (const auto &g : G) {
        const uint32_t m = (uint32_t)g.size();
        buf.clear(); off.assign(1, 0); lonely.clear(); order.resize(m);
        for (uint32_t t = 0; t < m; t++) { lonely.push_back(append_key(pk, kind, g[t], buf)); off.push_back((uint32_t)buf.size()); order[t] = t; }
        const int64_t *B = buf.data();
        auto cmp3 = [&](uint32_t a, uint32_t b) {              // lexicographic, shorter prefix first (Vec<T>'s Ord)
            const uint32_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b], l = std::min(la, lb);
            for (uint32_t i = 0; i < l; i++) { const int64_t x = B[off[a] + i], y = B[off[b] + i]; if (x != y) return x < y ? -1 : 1; }
            return la == lb ? 0 : (la < lb ? -1 : 1);
        };
        // Only equality of keys matters for the partition (the order of the groups does not): sort by a 64-bit hash of the
        // key, then by index, and fall back to the full comparison inside a run of equal hashes that holds different keys.
        hk.resize(m);
        for (uint32_t t = 0; t < m; t++) {
            uint64_t h = 0x9E3779B97F4A7C15ull ^ (off[t + 1] - off[t]);
            for (uint32_t i = off[t]; i < off[t + 1]; i++) { h ^= (uint64_t)B[i]; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29; }
            hk[t] = { h, ((uint64_t)g[t] << 32) | t };
        }
        std::sort(hk.begin(), hk.end());
        for (uint32_t i = 0; i < m;) {
            uint32_t j = i + 1;
            bool mixed = false;
            while (j < m && hk[j].first == hk[i].first) { mixed = mixed || cmp3((uint32_t)hk[i].second, (uint32_t)hk[j].second) != 0; j++; }
            for (uint32_t k = i; k < j; k++) order[k] = (uint32_t)hk[k].second;
            if (mixed) std::sort(order.begin() + i, order.begin() + j, [&](uint32_t a, uint32_t b) { const int c = cmp3(a, b); return c ? c < 0 : g[a] < g[b]; });
            i = j;
        }
        for (uint32_t i = 0; i < m;) {
            uint32_t j = i + 1;
            while (j < m && cmp3(order[i], order[j]) == 0) j++;
            if (lonely[order[i]]) for (uint32_t k = i; k < j; k++) out.push_back({ g[order[k]] });
            else { out.emplace_back(); out.back().reserve(j - i); for (uint32_t k = i; k < j; k++) out.back().push_back(g[order[k]]); }
            i = j;
        }
    }
    G.swap(out);
}

struct DevMem {                    // stream-ordered allocations from the device's default pool (kept warm across calls)
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit DevMem(cudaStream_t s) : st(s) {}
    ~DevMem() { for (void *p : ptrs) cudaFreeAsync(p, st); }
    template <typename T> cudaError_t take(T *&dst, size_t count) {
        void *p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T) + 64, st);
        if (e != cudaSuccess) return e;
        ptrs.push_back(p);
        dst = (T *)p;
        return cudaSuccess;
    }
    template <typename T> cudaError_t up(const T *&dst, const T *src, size_t count) {
        T *p = nullptr;
        cudaError_t e = take(p, count);
        if (e != cudaSuccess) return e;
        dst = p;
        return count ? cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st) : cudaSuccess;
    }
};

struct StreamBox { cudaStream_t s = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr; ~StreamBox() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } };

double now_ms() { using namespace std::chrono; return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count(); }

int upload_pack(DevMem &M, const egroup_pack_t *pk, GDev &d) {
    const uint32_t n = pk->n_clono;
    GCK(M.up(d.clono_exact_off, pk->clono_exact_off, (size_t)n + 1));
    GCK(M.up(d.clono_exact, pk->clono_exact, pk->clono_exact_off[n]));
    GCK(M.up(d.exact_chain_off, pk->exact_chain_off, (size_t)pk->n_exact + 1));
    GCK(M.up(d.left, pk->left, pk->n_chain));
    GCK(M.up(d.seq_off, pk->seq_off, pk->n_chain)); GCK(M.up(d.seq_len, pk->seq_len, pk->n_chain));
    GCK(M.up(d.cdr3_start, pk->cdr3_start, pk->n_chain));
    GCK(M.up(d.cdr3_aa_off, pk->cdr3_aa_off, pk->n_chain)); GCK(M.up(d.cdr3_aa_len, pk->cdr3_aa_len, pk->n_chain));
    GCK(M.up(d.seq_arena, pk->seq_arena, pk->seq_arena_bytes)); GCK(M.up(d.aa_arena, pk->aa_arena, pk->aa_arena_bytes));
    return 0;
}

int check_group_pack(const egroup_pack_t *pk) {
    if (pk->abi_version != EGROUP_ABI_VERSION) return gfail(-1, "abi_version mismatch");
    if (pk->n_clono && (!pk->clono_exact_off || !pk->clono_exact || !pk->exact_chain_off || !pk->left || !pk->seq_off || !pk->seq_len || !pk->seq_del_len ||
                        !pk->cdr3_start || !pk->cdr3_aa_off || !pk->cdr3_aa_len || !pk->seq_arena || !pk->aa_arena || !pk->col_off))
        return gfail(-1, "null array in pack");
    return 0;
}

struct AsymOwner { std::vector<uint64_t> off; std::vector<uint32_t> member; std::vector<double> dist; };

}  // namespace

extern "C" const char *egroup_last_error(void) { return g_err.c_str(); }

extern "C" void egroup_opts_default(egroup_opts_t *o) {
    memset(o, 0, sizeof(*o));
    o->heavy_pc = o->light_pc = o->cdr3_heavy_pc_hf = o->aa_heavy_pc = o->aa_light_pc = o->cdr3_heavy_pc = o->cdr3_light_pc = o->cdr3_aa_heavy_pc = o->cdr3_aa_light_pc = -1.0;
}

extern "C" int egroup_run(const egroup_pack_t *pk, const egroup_opts_t *op, int device, uint32_t *group_of, egroup_stats_t *stats) {
    if (!pk || !op || (pk->n_clono && !group_of)) return gfail(-1, "null argument");
    if (int rc = check_group_pack(pk)) return rc;
    const uint32_t n = pk->n_clono;
    if (op->cdr3_heavy_pc_hf >= 0.0 && !op->hf_matrix) return gfail(-1, "cdr3_heavy_pc_hf without its penalty matrix");
    const double t0 = now_ms();
    egroup_stats_t st;
    memset(&st, 0, sizeof(st));
    Groups G(1);
    for (uint32_t i = 0; i < n; i++) G[0].push_back(i);
    if (n == 0) G.clear();
    const bool dbg = getenv("EGROUP_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[egroup] init %.3f ms\n", now_ms() - t0);
    if (op->vj_refname) key_pass(pk, K_VJ, G);
    if (op->vdj_refname) key_pass(pk, K_VDJ, G);
    if (op->v_heavy_refname) key_pass(pk, K_VH, G);
    if (op->vj_heavy_refname) key_pass(pk, K_VJH, G);
    if (op->vdj_heavy_refname) key_pass(pk, K_VDJH, G);
    if (op->vj_len) key_pass(pk, K_VJLEN, G);
    if (op->cdr3_len) key_pass(pk, K_CDR3LEN, G);
    if (op->cdr3_heavy_len) key_pass(pk, K_CDR3HLEN, G);
    if (op->cdr3_light_len) key_pass(pk, K_CDR3LLEN, G);
    if (dbg) fprintf(stderr, "[egroup] key partitions done %.3f ms, %zu groups\n", now_ms() - t0, G.size());
    const struct { double pc; int kind, left; } passes[] = {
        { op->heavy_pc, D_SEQ, 1 }, { op->light_pc, D_SEQ, 0 }, { op->cdr3_heavy_pc_hf, D_HF, 1 }, { op->aa_heavy_pc, D_AA, 1 }, { op->aa_light_pc, D_AA, 0 },
        { op->cdr3_heavy_pc, D_CDR3, 1 }, { op->cdr3_light_pc, D_CDR3, 0 }, { op->cdr3_aa_heavy_pc, D_CDR3AA, 1 }, { op->cdr3_aa_light_pc, D_CDR3AA, 0 } };
    bool any_dist = false;
    for (const auto &p : passes) any_dist = any_dist || p.pc >= 0.0;
    if (any_dist && n) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) { cudaGetLastError(); return gfail(-3, "no usable CUDA device (the distance passes have no CPU fallback)"); }
        GCK(cudaSetDevice(device));
        StreamBox box;                                               // destroyed after the allocations below are returned to the pool
        GCK(cudaStreamCreateWithFlags(&box.s, cudaStreamNonBlocking));
        cudaStream_t s = box.s;
        DevMem M(s);
        GDev d;
        int sm = 148;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
        if (int rc = upload_pack(M, pk, d)) return rc;
        double penalty_h[27 * 27];
        memset(penalty_h, 0, sizeof(penalty_h));
        if (op->cdr3_heavy_pc_hf >= 0.0) {                                   // grouper.rs:463-472
            static const char *aa = "ACDEFGHIKLMNPQRSTVWY";
            for (int i1 = 0; i1 < 20; i1++)
                for (int i2 = 0; i2 < 20; i2++) penalty_h[(aa[i1] - 'A') * 27 + (aa[i2] - 'A')] = op->hf_matrix[i1 * 20 + i2];
        }
        const double *d_penalty = nullptr;
        GCK(M.up(d_penalty, penalty_h, 27 * 27));
        GCounters *d_ctr = nullptr;
        GCK(M.take(d_ctr, 1));
        GCK(cudaMemsetAsync(d_ctr, 0, sizeof(GCounters), s));
        if (pk->seq_arena_bytes) { k_validate_nt<<<sm * 4, 256, 0, s>>>(d.seq_arena, pk->seq_arena_bytes, d_ctr); st.gpu_launches++; }
        {   // keep the pool's memory between calls
            cudaMemPool_t pool; unsigned long long keep = ~0ull;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        GCK(cudaEventCreate(&box.e0)); GCK(cudaEventCreate(&box.e1));
        cudaEvent_t e0 = box.e0, e1 = box.e1;
        GCK(cudaEventRecord(e0, s));
        for (const auto &p : passes) {
            if (p.pc < 0.0) continue;
            // flatten the groups of this pass
            std::vector<uint32_t> members, grp_off;
            std::vector<unsigned long long> pair_off;
            unsigned long long np = 0;
            for (const auto &g : G) { grp_off.push_back((uint32_t)members.size()); pair_off.push_back(np); members.insert(members.end(), g.begin(), g.end()); np += (unsigned long long)g.size() * (g.size() - 1) / 2; }
            const uint32_t ng = (uint32_t)G.size(), nm = (uint32_t)members.size();
            st.pairs_candidate += np;
            std::vector<uint32_t> lab(nm);
            if (np) {
                DevMem P(s);
                const uint32_t *d_members = nullptr, *d_grp_off = nullptr;
                const unsigned long long *d_pair_off = nullptr;
                uint32_t *d_label = nullptr, *d_flat = nullptr;
                GCK(P.up(d_members, members.data(), nm)); GCK(P.up(d_grp_off, grp_off.data(), ng)); GCK(P.up(d_pair_off, pair_off.data(), ng));
                GCK(P.take(d_label, nm)); GCK(P.take(d_flat, nm));
                k_iota<<<(nm + 255) / 256, 256, 0, s>>>(d_label, nm);
                const unsigned long long want = (np + 127) / 128;
                const unsigned blocks = (unsigned)std::min<unsigned long long>(want, (unsigned long long)sm * 16);
                k_group_pairs<<<blocks, 128, 0, s>>>(d, p.kind, p.left, p.pc / 100.0, d_penalty, d_members, d_grp_off, d_pair_off, ng, np, d_label, d_ctr);
                k_flatten<<<(nm + 255) / 256, 256, 0, s>>>(d_label, nm, d_flat);
                st.gpu_launches += 3;
                GCK(cudaGetLastError());
                GCK(cudaMemcpyAsync(lab.data(), d_flat, 4 * (size_t)nm, cudaMemcpyDeviceToHost, s));
                GCK(cudaStreamSynchronize(s));
            } else for (uint32_t i = 0; i < nm; i++) lab[i] = i;
            // the orbits of this pass, listed the way the source lists them (grouper.rs:441-452): orbit(i) for i < #orbits
            Groups out;
            for (uint32_t gi = 0; gi < ng; gi++) {
                const uint32_t base = grp_off[gi], m = (uint32_t)G[gi].size();
                uint32_t nreps = 0;
                for (uint32_t i = 0; i < m; i++) nreps += lab[base + i] == base + i;          // labels are min members: a rep labels itself
                for (uint32_t i = 0; i < nreps; i++) {
                    out.emplace_back();
                    for (uint32_t j = 0; j < m; j++) if (lab[base + j] == lab[base + i]) out.back().push_back(G[gi][j]);
                }
            }
            G.swap(out);
        }
        GCK(cudaEventRecord(e1, s));
        GCounters hc;
        GCK(cudaMemcpyAsync(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, s));
        GCK(cudaStreamSynchronize(s));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        st.ms_device = ms; st.pairs_evaluated = hc.pairs_evaluated; st.chain_compares = hc.chain_compares;
        if (hc.err & ERR_BAD_BYTE) return gfail(-4, "a nucleotide sequence holds a byte outside ACGTN-");
        if (hc.err & ERR_TOO_LONG) return gfail(-4, "a sequence compared by a distance pass is longer than 512 symbols");
    }
    if (dbg) fprintf(stderr, "[egroup] distance passes done %.3f ms\n", now_ms() - t0);
    // grouper.rs:760-766: join consecutive members of every group; labels = smallest member
    std::vector<uint32_t> parent(n);
    for (uint32_t i = 0; i < n; i++) parent[i] = i;
    auto find = [&](uint32_t a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    for (const auto &g : G)
        for (size_t i = 0; i + 1 < g.size(); i++) { uint32_t a = find(g[i]), b = find(g[i + 1]); if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; } }
    uint64_t ncls = 0;
    for (uint32_t i = 0; i < n; i++) { group_of[i] = find(i); ncls += group_of[i] == i; }
    st.n_groups = ncls;
    st.ms_total = now_ms() - t0;
    if (stats) *stats = st;
    return 0;
}

extern "C" void egroup_asym_free(egroup_asym_result_t *r) {
    if (!r) return;
    delete (AsymOwner *)r->owner;
    memset(r, 0, sizeof(*r));
}

extern "C" int egroup_run_asym(const egroup_pack_t *pk, const uint8_t *in_center, const egroup_asym_opts_t *op, int device,
                               egroup_asym_result_t *out, egroup_stats_t *stats) {
    if (!pk || !op || !out || (pk->n_clono && !in_center)) return gfail(-1, "null argument");
    if (int rc = check_group_pack(pk)) return rc;
    if (op->bound_kind < EGROUP_BOUND_TOP || op->bound_kind > EGROUP_BOUND_NONE) return gfail(-1, "bound_kind");
    memset(out, 0, sizeof(*out));
    const double t0 = now_ms();
    egroup_stats_t st;
    memset(&st, 0, sizeof(st));
    const uint32_t n = pk->n_clono;
    std::vector<uint32_t> center;
    for (uint32_t i = 0; i < n; i++) if (in_center[i]) center.push_back(i);             // grouper.rs:832-853
    const uint32_t nc = (uint32_t)center.size();
    AsymOwner *own = new AsymOwner();
    out->owner = own;
    struct Guard { egroup_asym_result_t *o; bool ok; ~Guard() { if (!ok) egroup_asym_free(o); } } guard{ out, false };
    // per center: the (index, distance) candidates the walk of grouper.rs:909-920 can reach, unsorted
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> cand(nc);                   // (distance, index)
    if (nc) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) { cudaGetLastError(); return gfail(-3, "no usable CUDA device (the distance matrix has no CPU fallback)"); }
        GCK(cudaSetDevice(device));
        StreamBox box;
        GCK(cudaStreamCreateWithFlags(&box.s, cudaStreamNonBlocking));
        cudaStream_t s = box.s;
        DevMem M(s);
        GDev d;
        int sm = 148;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
        if (int rc = upload_pack(M, pk, d)) return rc;
        const uint32_t *d_center = nullptr;
        GCK(M.up(d_center, center.data(), nc));
        GCounters *d_ctr = nullptr;
        GCK(M.take(d_ctr, 1));
        GCK(cudaMemsetAsync(d_ctr, 0, sizeof(GCounters), s));
        GCK(cudaEventCreate(&box.e0)); GCK(cudaEventCreate(&box.e1));
        GCK(cudaEventRecord(box.e0, s));
        st.pairs_candidate = (uint64_t)nc * n;
        if (op->bound_kind == EGROUP_BOUND_TOP) {
            const uint32_t need = (uint32_t)std::min<uint64_t>((uint64_t)op->top + 1, n);
            // one row of distances per resident CTA; the grid is cut so that the rows stay under 512 MB
            unsigned blocks = (unsigned)std::min<uint64_t>(nc, (uint64_t)sm * 4);
            blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(blocks, (512ull << 20) / (4ull * n)));
            uint32_t *d_row = nullptr;
            uint2 *d_out = nullptr;
            GCK(M.take(d_row, (size_t)blocks * n));
            GCK(M.take(d_out, (size_t)nc * need));
            k_asym_top<<<blocks, 256, 0, s>>>(d, d_center, nc, n, need, d_row, d_out, d_ctr);
            st.gpu_launches++;
            GCK(cudaGetLastError());
            std::vector<uint2> h((size_t)nc * need);
            GCK(cudaMemcpyAsync(h.data(), d_out, sizeof(uint2) * h.size(), cudaMemcpyDeviceToHost, s));
            GCK(cudaStreamSynchronize(s));
            for (uint32_t i = 0; i < nc; i++)
                for (uint32_t k = 0; k < need; k++) cand[i].push_back({ h[(size_t)i * need + k].y, h[(size_t)i * need + k].x });
        } else {
            // d <= D for integer d  <=>  d <= floor(D); no bound: everything (distances never exceed ASYM_INF)
            uint32_t dcut = ASYM_INF;
            bool none = false;
            if (op->bound_kind == EGROUP_BOUND_MAX) {
                if (!(op->max_dist >= 0.0)) none = true;
                else dcut = op->max_dist >= (double)ASYM_INF ? ASYM_INF : (uint32_t)std::floor(op->max_dist);
            }
            if (!none) {
                unsigned long long *d_n = nullptr;
                GCK(M.take(d_n, 1));
                unsigned long long cap = std::max<unsigned long long>(1ull << 20, 64ull * nc), got = 0;
                const unsigned long long total = (unsigned long long)nc * n;
                cap = std::min(cap, total);
                for (int attempt = 0; attempt < 2; attempt++) {
                    uint4 *d_out = nullptr;
                    GCK(M.take(d_out, cap));
                    GCK(cudaMemsetAsync(d_n, 0, 8, s));
                    GCK(cudaMemsetAsync(d_ctr, 0, sizeof(GCounters), s));      // a second attempt recounts
                    // enough (center, slice) work items to fill the GPU a few times over, slices of at least 256 clonotypes
                    const uint32_t slices = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)sm * 8 + nc - 1) / nc, ((uint64_t)n + 255) / 256));
                    const unsigned blocks = (unsigned)std::min<unsigned long long>((unsigned long long)nc * slices, (unsigned long long)sm * 8);
                    k_asym_max<<<blocks, 256, 0, s>>>(d, d_center, nc, n, slices, dcut, d_out, cap, d_n, d_ctr);
                    st.gpu_launches++;
                    GCK(cudaGetLastError());
                    GCK(cudaMemcpyAsync(&got, d_n, 8, cudaMemcpyDeviceToHost, s));
                    GCK(cudaStreamSynchronize(s));
                    if (got <= cap) {
                        std::vector<uint4> h(got);
                        GCK(cudaMemcpyAsync(h.data(), d_out, sizeof(uint4) * got, cudaMemcpyDeviceToHost, s));
                        GCK(cudaStreamSynchronize(s));
                        for (const uint4 &e : h) cand[e.x].push_back({ e.z, e.y });
                        break;
                    }
                    cap = got;                                   // the list was too small: once more with the size it asked for
                }
            }
        }
        GCK(cudaEventRecord(box.e1, s));
        GCounters hc;
        GCK(cudaMemcpyAsync(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, s));
        GCK(cudaStreamSynchronize(s));
        float ms = 0;
        cudaEventElapsedTime(&ms, box.e0, box.e1);
        st.ms_device = ms; st.pairs_evaluated = hc.pairs_evaluated; st.chain_compares = hc.chain_compares;
        if (hc.err & ERR_TOO_LONG) return gfail(-4, "a CDR3 amino acid sequence longer than 512");
    }
    // grouper.rs:891-925
    own->off.push_back(0);
    for (uint32_t i = 0; i < nc; i++) {
        auto &id = cand[i];
        std::sort(id.begin(), id.end());                                                  // (distance, index), :908
        const size_t start = own->member.size();
        own->member.push_back(center[i]); own->dist.push_back(0.0);                       // :903
        for (size_t p = 0; p < id.size(); p++) {
            if (op->bound_kind == EGROUP_BOUND_TOP && p > op->top) break;
            if (op->bound_kind == EGROUP_BOUND_MAX && (double)id[p].first > op->max_dist) break;
            if (id[p].second != center[i]) { own->member.push_back(id[p].second); own->dist.push_back((double)id[p].first); }
        }
        if (own->member.size() - start >= op->min_group) own->off.push_back(own->member.size());
        else { own->member.resize(start); own->dist.resize(start); }
    }
    out->n_groups = own->off.size() - 1; out->group_off = own->off.data(); out->member = own->member.data(); out->dist = own->dist.data();
    st.n_groups = out->n_groups;
    st.ms_total = now_ms() - t0;
    if (stats) *stats = st;
    guard.ok = true;
    return 0;
}
