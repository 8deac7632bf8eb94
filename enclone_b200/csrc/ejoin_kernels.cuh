// ejoin_kernels.cuh — the sm_100a kernels of the join path.  See ejoin_dev.cuh for layouts.
#pragma once
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include "ejoin_dev.cuh"


__device__ __forceinline__ uint32_t uf_find(uint32_t *p, uint32_t x);
__device__ __forceinline__ void uf_union(uint32_t *p, uint32_t a, uint32_t b);

struct MaxOp {
    template <typename T> __host__ __device__ __forceinline__ T operator()(const T &a, const T &b) const { return a > b ? a : b; }
};

// ---------------------------------------------------------------------------------------------
// K0  bucket heads: a bucket is a maximal run of equal lens.
// ---------------------------------------------------------------------------------------------
__global__ void k_bucket_heads(DevIn in, uint32_t *head_pos) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= in.n_info) return;
    bool head = k == 0 || in.len[0][k] != in.len[0][k - 1] || in.len[1][k] != in.len[1][k - 1];
    head_pos[k] = head ? k : 0u;
}

// K1  sort keys + validity + presence bitmaps + candidate-pair count.
// Entries that can never join (one chain; exact subclonotype without 2 or 3 chains:
// help1.rs:270-276) get the key ~0 and drop out of every comparison group.
__global__ void k_entry_keys(DevIn in, const uint32_t *bucket_start, unsigned long long *keys, uint32_t *vals,
                             uint32_t *presL, uint32_t *presC, DevCounters *ctr) {
    // grid-stride: a few hundred blocks, each ending with ONE pair of atomics on the two global tallies (one block per 256
    // entries meant 10,000 same-address atomics in a row: 0.1 ms at 1.27 M entries, four times the rest of the kernel)
    unsigned long long pairs = 0, heads = 0;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < in.n_info; k += gridDim.x * blockDim.x) {
        uint32_t l0 = in.len[0][k], l1 = in.len[1][k], c0 = in.cdr3_len[0][k], c1 = in.cdr3_len[1][k];
        uint32_t e = in.exact_id[k];
        uint32_t nch = e < in.n_exact ? in.nchains[e] : 0;
        if (e >= in.n_exact) atomicOr(&ctr->err, ERR_BAD_EXACT);
        bool valid = l1 > 0 && nch >= 2 && nch <= 3;
        unsigned long long key = ~0ull;
        if (valid) {
            if (c0 > 255 || c1 > 255 || c0 + c1 > C3_MAX_TOTAL) { atomicOr(&ctr->err, ERR_CDR3_TOO_LONG); valid = false; }
            else if (l0 + l1 >= 8192) { atomicOr(&ctr->err, ERR_LEN_TOO_LARGE); valid = false; }
        }
        if (valid) {
            // bucket | cdr3len0:8 | cdr3len1:8, the bucket named by the index of its first entry (info is in bucket order and the
            // sort is stable, so this orders exactly as len0 | len1 | ... would, in 16 + log2(n) key bits instead of 42: one
            // radix pass fewer; two separate runs of equal lens stay two buckets, as in the reference's bucket scan)
            key = ((unsigned long long)bucket_start[k] << 16) | ((unsigned long long)c0 << 8) | c1;
            uint32_t L = l0 + l1, cc = (c0 << 8) | c1;
            // test before set: after the first hit per value these are plain cached loads, not serialized atomics.  (Measured: reading the
            // bitmaps through L2 with __ldcg instead, or twice as many blocks, each made the keys stage 0.02 ms slower.)
            if (!(presL[L >> 5] & (1u << (L & 31)))) atomicOr(&presL[L >> 5], 1u << (L & 31));
            if (!(presC[cc >> 5] & (1u << (cc & 31)))) atomicOr(&presC[cc >> 5], 1u << (cc & 31));
        }
        keys[k] = key;
        vals[k] = k;
        pairs += k - bucket_start[k];
        heads += bucket_start[k] == k;
    }
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    unsigned long long s = BR(tmp).Sum(pairs);
    __syncthreads();
    unsigned long long h = BR(tmp).Sum(heads);
    if (threadIdx.x == 0) { atomicAdd(&ctr->pairs_candidate, s); atomicAdd(&ctr->n_buckets, h); }
}

// K2  group table from the run-length encoded sorted keys.
struct GroupVals { unsigned long long v[6]; uint32_t W, cdm, T; bool valid; };   // entries, work items, c3 words, pairs, tier-1 bytes, extra slot rows
__device__ __forceinline__ GroupVals group_vals(unsigned long long key, unsigned long long n, const uint16_t *__restrict__ cdmax_tab, uint32_t seg_min) {
    GroupVals o;
#pragma unroll
    for (int f = 0; f < 6; f++) o.v[f] = 0;
    o.W = 0; o.cdm = 0xFFFF; o.T = 0; o.valid = key != ~0ull;
    if (!o.valid) return o;
    const uint32_t total = (uint32_t)((key >> 8) & 0xFF) + (uint32_t)(key & 0xFF);
    uint32_t W = (total + 31) / 32;
    if (W == 0) W = 1;
    const uint32_t cdm = cdmax_tab[total];
    const uint32_t T = (uint32_t)((n + T1_TILE - 1) / T1_TILE);
    o.W = W; o.cdm = cdm; o.T = T;
    o.v[0] = n;
    if (n >= 2 && cdm != 0xFFFF) {
        const uint32_t S = seg_of_tiles(T, seg_min), Q = (T + S - 1) / S;
        const unsigned long long items = seg_item_off(T, S, Q);       // = Q*T - S*Q*(Q-1)/2
        o.v[1] = items;
        o.v[3] = n * (n - 1) / 2;
        o.v[4] = o.v[3] * 16ull * W;
        o.v[5] = items - T;
    }
    o.v[2] = (n * 2 * W + 3) & ~3ull;
    return o;
}
// Three kernels around three device-wide exclusive sums (CUB decoupled look-back): the number of runs is only known on
// the device, so the scans run over n slots (runs <= n) with zeros behind the last run.
//   sumA = entries | work items << 32   (both totals < 2^32),   sumB = c3 words,   sumC = extra slot rows
__global__ void __launch_bounds__(256) k_group_vals(const unsigned long long *__restrict__ uniq, const uint32_t *__restrict__ counts,
                                                    const uint32_t *num_runs, uint32_t n, const uint16_t *__restrict__ cdmax_tab,
                                                    unsigned long long *__restrict__ sumA, unsigned long long *__restrict__ sumB,
                                                    uint32_t *__restrict__ sumC, DevCounters *ctr, uint32_t seg_min) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nr = *num_runs;
    unsigned long long pairs = 0, t1b = 0, ng = 0;
    if (r < n) {
        unsigned long long a = 0, b = 0;
        uint32_t x = 0;
        if (r < nr) {
            const GroupVals gv = group_vals(uniq[r], counts[r], cdmax_tab, seg_min);
            a = gv.v[0] | (gv.v[1] << 32); b = gv.v[2]; x = (uint32_t)gv.v[5];
            pairs = gv.v[3]; t1b = gv.v[4];
            if (gv.valid) ng = r + 1;
            if (gv.T > 65535u) atomicOr(&ctr->err, ERR_GROUP_TOO_LARGE);      // GroupInfo::tiles is 16 bits
        }
        sumA[r] = a; sumB[r] = b; sumC[r] = x;
    }
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    pairs = BR(tmp).Sum(pairs);
    __syncthreads();
    t1b = BR(tmp).Sum(t1b);
    __syncthreads();
    ng = BR(tmp).Reduce(ng, MaxOp());
    if (threadIdx.x == 0) {
        if (pairs) { atomicAdd(&ctr->pairs_compared, pairs); atomicAdd(&ctr->t1_alg_bytes, t1b); }
        if (ng) atomicMax(&ctr->n_groups, ng);         // number of groups = number of valid runs (they precede the ~0 run)
    }
}
__global__ void __launch_bounds__(256) k_group_build(const unsigned long long *__restrict__ uniq, const uint32_t *__restrict__ counts,
                                                     const uint32_t *num_runs, const uint16_t *__restrict__ cdmax_tab,
                                                     const unsigned long long *__restrict__ sumA, const unsigned long long *__restrict__ sumB,
                                                     const uint32_t *__restrict__ sumC, GroupInfo *__restrict__ groups, DevCounters *ctr, uint32_t seg_min,
                                                     uint32_t mma_min_tiles, unsigned long long i8_cap_tiles) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nr = *num_runs;
    if (r >= nr) return;
    const GroupVals gv = group_vals(uniq[r], counts[r], cdmax_tab, seg_min);
    const unsigned long long a = sumA[r], b = sumB[r];
    const uint32_t x = sumC[r];
    if (gv.valid) {
        GroupInfo g;
        g.start = (uint32_t)(a & 0xFFFFFFFFull);
        g.n = counts[r];
        g.c3base = (uint32_t)b;
        g.work_off = (uint32_t)(a >> 32);
        g.xoff = x;
        g.W = (uint16_t)gv.W; g.cdmax = (uint16_t)gv.cdm; g.tiles = (uint16_t)gv.T; g.seg = (uint16_t)seg_of_tiles(gv.T, seg_min); g.mma = MMA_NONE;
        {   // large groups of short CDR3s: tier 1 on the tensor cores, if int8 tile slots are left (none left: the popcount sweep takes it)
            const uint32_t total = (uint32_t)((uniq[r] >> 8) & 0xFF) + (uint32_t)(uniq[r] & 0xFF);
            if (mma_min_tiles && gv.T >= mma_min_tiles && gv.T <= 65535u && total >= 1 && total <= MMA_MAX_BASES && g.n >= 2 && gv.cdm != 0xFFFF && 3 * total > 4 * gv.cdm) {
                const uint32_t slot = atomicAdd(&ctr->i8_tiles, gv.T);
                if ((unsigned long long)slot + gv.T <= i8_cap_tiles && slot + gv.T < (1u << 23)) {
                    g.mma = (total << 23) | slot;
                    atomicAdd(&ctr->mma_pairs, gv.v[3]);
                    atomicAdd(&ctr->mma_macs, (unsigned long long)gv.T * (gv.T + 1) / 2 * T1_TILE * T1_TILE * (MMA_KSTEPS(total) * 32));
                    atomicAdd(&ctr->t1_alg_bytes, 0ull - gv.v[4]);            // not popcount work (k_group_vals counted every group)
                }
            }
        }
        groups[r] = g;               // valid runs precede the ~0 run, so r is the group id
    }
    if (r == nr - 1) {               // totals = exclusive prefix of the last run + its own values
        ctr->n_valid = (a & 0xFFFFFFFFull) + gv.v[0];
        ctr->n_work = (a >> 32) + gv.v[1];
        ctr->c3_words = b + gv.v[2];
        ctr->n_xrows = (unsigned long long)x + gv.v[5];
    }
}

// Work items sorted by cost, longest first (counting sort over 64 cost classes): the sweep's CTAs take
// items from one atomic counter, and with the long items in front the tail of the kernel is made of
// the short ones.  In stride mode (a giant bucket shared by several ranks) only the rows this rank
// owns enter the list: row r of the global numbering belongs to rank r % stride.
__device__ __forceinline__ bool item_decode(const GroupInfo *__restrict__ groups, uint32_t ngroups, unsigned long long item, uint32_t stride,
                                            uint32_t phase, uint32_t *g_out, uint32_t *q_out, uint32_t *ib_out, uint32_t *cls_out) {
    uint32_t lo = 0, hi = ngroups;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (groups[mid].work_off <= item) lo = mid; else hi = mid; }
    const GroupInfo g = groups[lo];
    const uint32_t T = g.tiles, S = g.seg;
    uint32_t rem = (uint32_t)(item - g.work_off), q = 0;
    while (rem >= T - q * S) { rem -= T - q * S; q++; }
    const uint32_t ib = T - 1 - rem;
    *g_out = lo; *q_out = q; *ib_out = ib;
    if (stride > 1 && (g.work_off + (T - 1 - ib)) % stride != phase) return false;
    // cost ~ half tile steps (the diagonal tile counts half) x rows in the B tile x words per row
    const uint32_t hi_t = min((q + 1) * S - 1, ib);
    const uint32_t half = 2 * (hi_t - q * S + 1) - (hi_t == ib ? 1u : 0u);
    const uint32_t rowsB = min((uint32_t)T1_TILE, g.n - ib * T1_TILE);
    const uint32_t cost = (half * g.W * ((rowsB + 31) >> 5) + 7) >> 3;          // 16 half steps x W=3 x 4 warps -> 24
    *cls_out = min(cost, 63u);
    return true;
}
__global__ void __launch_bounds__(256) k_item_hist(const GroupInfo *__restrict__ groups, DevCounters *ctr, uint8_t *__restrict__ item_cls,
                                                   uint32_t item_cap, uint32_t stride, uint32_t phase) {
    __shared__ uint32_t hist[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t ngroups = (uint32_t)ctr->n_groups;
    const unsigned long long nwork = min(ctr->n_work, (unsigned long long)item_cap);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwork; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t g, q, ib, cls;
        const bool own = item_decode(groups, ngroups, i, stride, phase, &g, &q, &ib, &cls);
        item_cls[i] = own ? (uint8_t)cls : (uint8_t)255;
        if (own) atomicAdd(&hist[cls], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 64 && hist[threadIdx.x]) atomicAdd(&ctr->cls_count[threadIdx.x], hist[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_item_scatter(const GroupInfo *__restrict__ groups, DevCounters *ctr, const uint8_t *__restrict__ item_cls,
                                                      uint2 *__restrict__ order, uint32_t item_cap, uint32_t stride, uint32_t phase) {
    __shared__ uint32_t base[64];
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int cl = 63; cl >= 0; cl--) { base[cl] = run; run += ctr->cls_count[cl]; }
        if (blockIdx.x == 0) ctr->n_items = run;
    }
    __syncthreads();
    const uint32_t ngroups = (uint32_t)ctr->n_groups;
    const unsigned long long nwork = min(ctr->n_work, (unsigned long long)item_cap);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwork; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t cl = item_cls[i];
        if (cl == 255) continue;
        uint32_t g, q, ib, cls;
        item_decode(groups, ngroups, i, stride, phase, &g, &q, &ib, &cls);
        const uint32_t pos = base[cl] + atomicAdd(&ctr->cls_cursor[cl], 1u);
        order[pos] = make_uint2(g, (q << 16) | ib);
    }
}

// head flag of the sorted keys (1 where a new comparison group starts); its inclusive sum - 1 is
// the group id of the entry (RLE run index)
__global__ void k_sorted_heads(const unsigned long long *skeys, uint32_t n, uint32_t *flag) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) flag[s] = (s == 0 || skeys[s] != skeys[s - 1]) ? 1u : 0u;
}

// per sorted entry: t2 row size in 16-byte units = 3 * (ceil(len0/128) + ceil(len1/128))
__global__ void __launch_bounds__(256) k_t2_sizes(DevIn in, const uint32_t *perm, DevCounters *ctr, uint32_t *sizes) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = 0;
    if (s < in.n_info) {
        if (s < ctr->n_valid) {
            uint32_t k = perm[s];
            v = 3u * (((uint32_t)in.len[0][k] + 127) / 128 + ((uint32_t)in.len[1][k] + 127) / 128);
        }
        sizes[s] = v;
    }
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    unsigned long long tot = BR(tmp).Sum((unsigned long long)v);
    if (threadIdx.x == 0 && tot) atomicAdd(&ctr->t2_units, tot);
}

// ---------------------------------------------------------------------------------------------
// K3  pack: builds the c3 rows, the t2 rows and the meta records straight from the caller's
// ASCII bytes (CloneInfo.tigs / cdr3s / vs / js), in three kernels:
//   k_meta     one thread per entry: ids, lengths, window bounds, flags            -> meta
//   k_pack_c3  one thread per 32-base word of the concatenated CDR3s               -> c3 rows
//   k_pack_t2  one half warp per entry, one LANE per 32-base plane word: 32 bases are read with three
//              16-byte loads per sequence and reduced in registers with SWAR byte tricks
//              (no cross-lane traffic); H/L/M words go straight to the 128-bit rows       -> t2 rows
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t low_bytes(int nbytes) {       // 0xFF in the low nbytes (clamped to 0..4) bytes
    nbytes = max(0, min(4, nbytes));
    return (uint32_t)~(~0ull << (8 * nbytes));
}
__device__ __forceinline__ uint32_t low_bits(int nbits) {         // low nbits (clamped to 0..32) bits set
    nbits = max(0, min(32, nbits));
    return (uint32_t)~(~0ull << nbits);
}
__device__ __forceinline__ uint32_t nibble_of(uint32_t y) {       // bit 0 of byte i -> bit i
    return ((y & 0x01010101u) * 0x01020408u) >> 24 & 0xFu;
}
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t d) {   // bit 0 of each byte = (byte != 0)
    return ((d | ((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) >> 7) & 0x01010101u;
}
__device__ __forceinline__ uint32_t not_acgt(uint32_t x) {        // nonzero iff some byte is not one of A,C,G,T
    // rebuild the byte each 2-bit code (b>>1)&3 stands for with one PRMT and compare
    const uint32_t y = (x >> 1) & 0x03030303u;
    const uint32_t t = y | (y >> 4);
    const uint32_t sel = (t & 0xFFu) | ((t >> 8) & 0xFF00u);      // four code nibbles
    return __byte_perm(0x47544341u /* 'A','C','T','G' for codes 0,1,2,3 */, 0u, sel) ^ x;
}
// 32 bytes starting at an arbitrary address, from three aligned 16-byte loads (the arenas carry
// 64 bytes of slack).  out[i] holds bytes 4i..4i+3.
__device__ __forceinline__ void load32u(const uint8_t *p, uint32_t (&out)[8]) {
    const uintptr_t a = (uintptr_t)p;
    const uint4 *q = (const uint4 *)(a & ~(uintptr_t)15);
    const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
    const uint32_t w[12] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w };
    const uint32_t sh = 8u * (uint32_t)(a & 3);
    switch ((a >> 2) & 3) {
        case 0:
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = __funnelshift_r(w[i], w[i + 1], sh);
            break;
        case 1:
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = __funnelshift_r(w[i + 1], w[i + 2], sh);
            break;
        case 2:
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = __funnelshift_r(w[i + 2], w[i + 3], sh);
            break;
        default:
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = __funnelshift_r(w[i + 3], w[i + 4], sh);
            break;
    }
}
// plane words of 32 bytes: code = (byte >> 1) & 3, i.e. A,C,G,T -> 0,1,3,2 (any injective code
// works: only equality is ever used); H = bit 1 of the code, L = bit 0.
__device__ __forceinline__ void planes32(const uint32_t (&x)[8], uint32_t &H, uint32_t &L) {
    H = 0; L = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { H |= nibble_of(x[i] >> 2) << (4 * i); L |= nibble_of(x[i] >> 1) << (4 * i); }
}

// 32 bases of a DnaString word (base i at bits 63-2i, 62-2i; A,C,G,T = 0..3) -> the plane words of this library:
// bit i of H / L = high / low bit of the code (byte >> 1) & 3 that the ASCII path uses (A,C,G,T -> 0,1,3,2 = g ^ (g >> 1)).
__device__ __forceinline__ uint32_t even_bits16(uint32_t v) {      // bits 0,2,4,.. of v -> bits 0..15
    v &= 0x55555555u;
    v = (v | (v >> 1)) & 0x33333333u;
    v = (v | (v >> 2)) & 0x0F0F0F0Fu;
    v = (v | (v >> 4)) & 0x00FF00FFu;
    return (v | (v >> 8)) & 0xFFFFu;
}
__device__ __forceinline__ void planes_of_2bit(unsigned long long x, uint32_t &H, uint32_t &L) {
    // after the bit reversal base i sits at bits 2i (code bit 1) and 2i+1 (code bit 0)
    const unsigned long long r = __brevll(x);
    const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
    const uint32_t b1 = even_bits16(lo) | (even_bits16(hi) << 16), b0 = even_bits16(lo >> 1) | (even_bits16(hi >> 1) << 16);
    H = b1; L = b1 ^ b0;
}
// 16 consecutive bases of a chain as ASCII bytes, rebuilt from its H / L plane words (positions p0..p0+15, p0 % 16 == 0)
__device__ __forceinline__ void bytes_of_planes16(uint32_t H, uint32_t L, int half, uint32_t (&out)[4]) {
    const uint32_t h = (H >> (16 * half)) & 0xFFFFu, l = (L >> (16 * half)) & 0xFFFFu;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t hn = (h >> (4 * i)) & 0xFu, ln = (l >> (4 * i)) & 0xFu;
        // selector nibble j = 2 * hbit_j + lbit_j
        const uint32_t sh = (hn & 1u) | ((hn & 2u) << 3) | ((hn & 4u) << 6) | ((hn & 8u) << 9);
        const uint32_t sl = (ln & 1u) | ((ln & 2u) << 3) | ((ln & 4u) << 6) | ((ln & 8u) << 9);
        out[i] = __byte_perm(0x47544341u /* codes 0,1,2,3 -> 'A','C','T','G' */, 0u, (sh << 1) | sl);
    }
}

__global__ void __launch_bounds__(256) k_meta(DevIn in, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ gid_incl,
                                              const GroupInfo *__restrict__ groups, const uint32_t *__restrict__ t2off, DevCounters *ctr,
                                              EntryMeta *__restrict__ meta, uint32_t *__restrict__ k_to_s, const uint8_t *__restrict__ needed,
                                              int ref_v_trim, int ref_j_trim, uint32_t *__restrict__ asc_list) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (uint32_t)ctr->n_valid) return;
    const uint32_t k = perm[s];
    // input validation for every entry, scored or not
    if (in.vs_seq[0][k] >= in.n_refseq || in.js_seq[0][k] >= in.n_refseq || in.vs_seq[1][k] >= in.n_refseq || in.js_seq[1][k] >= in.n_refseq)
        atomicOr(&ctr->err, ERR_BAD_REFSEQ);
    if (!needed[s]) return;                       // never scored: no record needed
    const GroupInfo g = groups[gid_incl[s] - 1];
    EntryMeta m;
    memset(&m, 0, sizeof(m));
    const uint32_t e = in.exact_id[k];
    m.k = k; m.t2off = t2off[s]; m.c3off = g.c3base + (s - g.start) * 2 * g.W;
    m.donor_lo = in.donor_lo[e]; m.donor_hi = in.donor_hi[e];
    uint32_t flags = 0;
    for (int c = 0; c < 2; c++) {
        const int n = in.len[c][k];
        const uint32_t vi = in.vs_seq[c][k], ji = in.js_seq[c][k];
        const bool bad = vi >= in.n_refseq || ji >= in.n_refseq;
        if (bad) atomicOr(&ctr->err, ERR_BAD_REFSEQ);
        const int vl = bad ? 0 : (int)in.refseq_len[vi], jl = bad ? 0 : (int)in.refseq_len[ji];
        int vend = max(0, min(n, vl - ref_v_trim));
        const int jw = jl - ref_j_trim;                  // J window length
        int jstart = jw > 0 ? max(0, n - jw) : n;
        if (jstart < vend) flags |= (FLAG_GENERIC0 << c);  // overlapping windows: counted separately -> bytewise path
        if (in.has_del[c][k]) flags |= (FLAG_GENERIC0 << c);
        m.vref[c] = in.v_ref_id[c][k]; m.jref[c] = in.j_ref_id[c][k]; m.dref[c] = in.dref_id[c][k];
        m.vs[c] = vi; m.js[c] = ji; m.vname[c] = in.vname_id[c][k];
        m.len[c] = (uint16_t)n; m.cs[c] = in.cdr3_start[c][k]; m.cl[c] = in.cdr3_len[c][k];
        m.vend[c] = (uint16_t)vend; m.jstart[c] = (uint16_t)jstart;
    }
    m.c_light = in.c_ref_id_light[k];
    m.hcomp = in.hcomp[k]; m.fr1 = in.fr1[k]; m.cdr1 = in.cdr1[k]; m.fr2 = in.fr2[k]; m.cdr2 = in.cdr2[k]; m.fr3 = in.fr3[k];
    m.flags = (uint16_t)flags;
    meta[s] = m;
    k_to_s[k] = s;
    // 2-bit input: the few entries with an ASCII chain (deletion marker / non-ACGT byte) are listed for k_pack_t2
    if (in.tig2_src && (in.has_del[0][k] || in.has_del[1][k])) asc_list[atomicAdd(&ctr->n_asc, 1u)] = s;
}

// 4 threads per entry, thread j packs the c3 words j, j+4 (W <= 4 for CDR3 pairs up to 128 nt: 3 of 4 lanes busy).  The 32 bases of a word
// come from chain 0's CDR3, chain 1's, or both (the word that straddles the boundary).  Reads only the
// early-uploaded arrays (CDR3 strings and lengths), so it runs before the rest of the input has landed.
__global__ void __launch_bounds__(256) k_pack_c3(DevIn in, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ gid_incl,
                                                 const GroupInfo *__restrict__ groups, DevCounters *ctr, uint32_t *__restrict__ c3) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t s = t >> 2;
    if (s >= (uint32_t)ctr->n_valid) return;
    const uint32_t k = perm[s];
    const GroupInfo g = groups[gid_incl[s] - 1];
    const int cl0 = in.cdr3_len[0][k], total = cl0 + in.cdr3_len[1][k];
    const int W = g.W;
    const uint32_t c3off = g.c3base + (s - g.start) * 2 * W;
    {   // the entry's CDR3 data must lie inside the uploaded slice
        bool ok;
        if (in.cdr3_2bit) { const unsigned long long o = in.cdr3_2bit_off[k]; ok = o >= in.cdr3_lo && o + (unsigned)((total + 31) / 32) <= in.cdr3_hi; }
        else {
            const unsigned long long o0 = in.cdr3_off[0][k], o1 = in.cdr3_off[1][k];
            ok = o0 >= in.cdr3_lo && o0 + (unsigned)cl0 <= in.cdr3_hi && o1 >= in.cdr3_lo && o1 + (unsigned)(total - cl0) <= in.cdr3_hi;
        }
        if (!ok) { if ((t & 3) == 0) atomicOr(&ctr->err, ERR_ARENA_RANGE); for (uint32_t w = t & 3; (int)w < W; w += 4) { c3[c3off + w] = 0; c3[c3off + W + w] = 0; } return; }
    }
    for (uint32_t w = t & 3; (int)w < W; w += 4) {
    const int p0 = 32 * (int)w;
    const int nb = total - p0;                                 // valid bytes in this word
    if (in.cdr3_2bit) {                                        // the caller's word IS the row word, up to the bit order
        uint32_t H = 0, L = 0;
        if (nb > 0) planes_of_2bit(__ldg(in.cdr3_2bit + in.cdr3_2bit_off[k] + w), H, L);
        const uint32_t valid = low_bits(nb);
        c3[c3off + w] = H & valid;
        c3[c3off + W + w] = L & valid;
        continue;
    }
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = 0;
    if (p0 < cl0) {
        uint32_t a[8];
        load32u(in.cdr3_arena + in.cdr3_off[0][k] + p0, a);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = a[i] & low_bytes(cl0 - p0 - 4 * i);
    }
    if (p0 + 32 > cl0 && nb > 0) {
        uint32_t b[8];
        load32u(in.cdr3_arena + in.cdr3_off[1][k] + (p0 - cl0), b);      // may start a few bytes before the string: masked below
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] |= b[i] & ~low_bytes(cl0 - p0 - 4 * i);
    }
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] &= low_bytes(nb - 4 * i); bad |= not_acgt(x[i]) & low_bytes(nb - 4 * i); }
    if (bad) atomicOr(&ctr->err, ERR_CDR3_BAD_CHAR);
    uint32_t H, L;
    planes32(x, H, L);
    const uint32_t valid = low_bits(nb);
    c3[c3off + w] = H & valid;
    c3[c3off + W + w] = L & valid;
    }
}

// 32 bytes at byte offset `off` of a shared-memory staging buffer (16-byte aligned base)
__device__ __forceinline__ void load32s(const uint8_t *buf, uint32_t off, uint32_t (&out)[8]) {
    const uint4 *q = (const uint4 *)(buf + (off & ~15u));
    const uint4 v0 = q[0], v1 = q[1], v2 = q[2];
    const uint32_t w[12] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w };
    const uint32_t sh = 8u * (off & 3);
    switch ((off >> 2) & 3) {
        case 0:
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = __funnelshift_r(w[i], w[i + 1], sh);
            break;
        case 1:
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = __funnelshift_r(w[i + 1], w[i + 2], sh);
            break;
        case 2:
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = __funnelshift_r(w[i + 2], w[i + 3], sh);
            break;
        default:
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = __funnelshift_r(w[i + 3], w[i + 4], sh);
            break;
    }
}

// Germline table -> bit planes H, L (the 2-bit code) and N (byte is not one of ACGT), one block per sequence.
__global__ void __launch_bounds__(32) k_pack_ref(DevIn in) {
    const uint32_t i = blockIdx.x;
    if (i >= in.n_refseq) return;
    const int len = (int)in.refseq_len[i];
    const uint8_t *src = in.refseq_arena + in.refseq_off[i];
    const uint32_t TW = in.ref_plane_words, o = in.ref_plane_off[i];
    for (int w = (int)threadIdx.x; w * 32 < len; w += 32) {
        const int nb = len - 32 * w;
        uint32_t x[8];
        load32u(src + 32 * w, x);
        uint32_t N = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            x[q] &= low_bytes(nb - 4 * q);
            N |= ((nonzero_bytes(not_acgt(x[q]) & low_bytes(nb - 4 * q)) * 0x01020408u) >> 24 & 0xFu) << (4 * q);
        }
        uint32_t H, L;
        planes32(x, H, L);
        const uint32_t valid = low_bits(nb);
        in.ref_planes[o + w] = H & valid;
        in.ref_planes[TW + o + w] = L & valid;
        in.ref_planes[2 * TW + o + w] = N & valid;
    }
}

// [wm_lo, wm_hi): only the entries whose last V..J byte lies in this range of the caller's arena are
// packed by this launch (the arena is uploaded in chunks; see upload_range).  Each warp first copies
// its entry's bytes into shared memory with non-overlapping aligned 16-byte loads — in zero-copy mode
// in.tig_arena is the caller's pinned host buffer and these loads are the only PCIe traffic for the
// V..J bytes, issued for needed entries only — and then every lane reduces its 32-base word from there.
// EJOIN_F_ENTRY_RECS: one 64-byte record per entry -> the parallel arrays the kernels read (device to device, ~40 us at
// 1.27 M entries); 0xFFFF widens to EJOIN_NONE.
struct RecOut {
    uint64_t *tig_off[2]; uint8_t *has_del[2]; uint16_t *cdr3_start[2];
    uint32_t *v_ref_id[2], *j_ref_id[2], *dref_id[2], *vs_seq[2], *js_seq[2], *vname_id[2], *vuni_seq[2], *utr_seq[2];
    uint32_t *c_ref_id_light; uint16_t *hcomp, *fr1, *cdr1, *fr2, *cdr2, *fr3;
};
__device__ __forceinline__ uint32_t widen16(uint16_t v) { return v == 0xFFFFu ? EJOIN_NONE : (uint32_t)v; }
__global__ void __launch_bounds__(256) k_unpack_recs(const ejoin_entry_rec_t *__restrict__ rec, uint32_t n, RecOut o) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    union { ejoin_entry_rec_t r; uint4 q[4]; } u;
    const uint4 *src = (const uint4 *)(rec + k);
#pragma unroll
    for (int i = 0; i < 4; i++) u.q[i] = __ldg(src + i);
    const ejoin_entry_rec_t &r = u.r;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        o.tig_off[c][k] = (unsigned long long)r.tig_off[c];
        o.has_del[c][k] = r.has_del[c]; o.cdr3_start[c][k] = r.cdr3_start[c];
        o.v_ref_id[c][k] = widen16(r.v_ref_id[c]); o.j_ref_id[c][k] = widen16(r.j_ref_id[c]); o.dref_id[c][k] = widen16(r.dref_id[c]);
        o.vs_seq[c][k] = widen16(r.vs_seq[c]); o.js_seq[c][k] = widen16(r.js_seq[c]); o.vname_id[c][k] = widen16(r.vname_id[c]);
        o.vuni_seq[c][k] = widen16(r.vuni_seq[c]); o.utr_seq[c][k] = widen16(r.utr_seq[c]);
    }
    o.c_ref_id_light[k] = widen16(r.c_ref_id_light); o.hcomp[k] = r.hcomp;
    o.fr1[k] = r.fr1_start; o.cdr1[k] = r.cdr1_start; o.fr2[k] = r.fr2_start; o.cdr2[k] = r.cdr2_start; o.fr3[k] = r.fr3_start;
}

#define PACK_STAGE 544          // per chain: 512 bytes + alignment slack + the 48-byte window overrun
// one HALF warp per entry: a 2 x 350-base entry has 44 plane words = 3 passes of 16 lanes (92 % of the lane slots;
// a whole warp needed 2 passes of 32 = 69 %)
__device__ void pack_t2_entry(const DevIn &in, DevCounters *ctr, EntryMeta *__restrict__ meta, uint4 *__restrict__ t2, const uint8_t *__restrict__ needed,
                              unsigned long long wm_lo, unsigned long long wm_hi, uint32_t s, uint8_t (*stage)[2][PACK_STAGE]) {
    const uint32_t lane = threadIdx.x & 15, wib = threadIdx.x >> 4;
    const uint32_t hmask = 0xFFFFu << (threadIdx.x & 16);          // the lanes of this half warp
    if (s >= (uint32_t)ctr->n_valid) return;
    if (!needed[s]) return;                       // no CDR3-compatible partner: its rows are never read
    const EntryMeta &m = meta[s];
    const uint32_t k = m.k;
    const int n0 = m.len[0], n1 = m.len[1];
    // which chains come as ASCII bytes (all of them unless the pack is EJOIN_F_TIGS_2BIT; then only those with has_del)
    const bool asc[2] = { in.tig2_src == nullptr || in.has_del[0][k] != 0, in.tig2_src == nullptr || in.has_del[1][k] != 0 };
    if (!asc[0] && !asc[1]) return;               // both chains 2 bits per base: k_pack_t2_2bit packs this entry
    if (in.tig2_src == nullptr) {
        const unsigned long long e0 = in.tig_off[0][k] + (unsigned long long)n0, e1 = in.tig_off[1][k] + (unsigned long long)n1;
        const unsigned long long last = e0 > e1 ? e0 : e1;
        if (!(last > wm_lo && last <= wm_hi)) return;
    }
    {   // every chain must lie inside the uploaded slice of its arena
        bool ok = true;
        for (int c = 0; c < 2; c++) {
            const unsigned long long o = in.tig_off[c][k];
            const unsigned nn = (unsigned)(c ? n1 : n0);
            if (asc[c]) ok = ok && (nn == 0 || (o >= in.tig_lo && o + nn <= in.tig_hi));
            else ok = ok && (nn == 0 || (o >= in.tig2_lo && o + (nn + 31) / 32 <= in.tig2_hi));
        }
        if (!ok) { if (lane == 0) atomicOr(&ctr->err, ERR_ARENA_RANGE); return; }
    }
    // ---- stage the ASCII chains (chains longer than 512 bases are read in place)
    uint32_t soff[2] = { 0, 0 };
    bool staged[2] = { false, false };
    for (int c = 0; c < 2; c++) {
        if (!asc[c]) continue;
        const int n = c ? n1 : n0;
        const uint8_t *t = in.tig_src + in.tig_off[c][k];
        const uintptr_t a = (uintptr_t)t;
        soff[c] = (uint32_t)(a & 15);
        staged[c] = n <= 512;
        const int nchunks = ((int)soff[c] + n + 15) / 16;              // <= 33 when staged
        const uint4 *src = (const uint4 *)(a & ~(uintptr_t)15);
        // the mirror has the same 16-byte phase as the source, so chunk i lands at the same arena offset
        uint4 *mir = in.tig_copy ? (uint4 *)(((uintptr_t)(in.tig_copy + in.tig_off[c][k])) & ~(uintptr_t)15) : nullptr;
        if (staged[c]) {
            uint4 *dst = (uint4 *)stage[wib][c];
            for (int i = (int)lane; i < nchunks; i += 16) { const uint4 v = src[i]; dst[i] = v; if (mir) mir[i] = v; }
        } else if (mir) {
            for (int i = (int)lane; i < nchunks; i += 16) mir[i] = src[i];
        }
        if (lane == 0 && in.tig_copy) atomicAdd(&ctr->inplace_bytes[blockIdx.x & 31], 16ull * nchunks);
    }
    __syncwarp(hmask);
    const int nw0 = 4 * ((n0 + 127) / 128), nw = nw0 + 4 * ((n1 + 127) / 128);
    uint32_t *rows = (uint32_t *)(t2 + m.t2off);
    if (lane == 0) {
        // bytes in (ASCII: one per base; 2-bit: one u64 word per 32 bases), three plane words out per 32 bases
        const unsigned long long bin0 = asc[0] ? (unsigned long long)n0 : 8ull * ((n0 + 31) / 32), bin1 = asc[1] ? (unsigned long long)n1 : 8ull * ((n1 + 31) / 32);
        atomicAdd(&ctr->pack_bytes[blockIdx.x & 31], bin0 + bin1 + 12ull * nw);
        if (in.tig2_src) atomicAdd(&ctr->inplace_bytes[blockIdx.x & 31], (asc[0] ? 0ull : bin0) + (asc[1] ? 0ull : bin1));
    }
    int mcount0 = 0, mcount1 = 0;
    uint32_t weird0 = 0, weird1 = 0;
    for (int w = (int)lane; w < nw; w += 16) {
        const int c = w >= nw0;
        const int wl = c ? w - nw0 : w;                       // word index inside the chain
        const int n = c ? n1 : n0;
        const int p0 = wl * 32;
        uint32_t H = 0, L = 0, M = 0;
        if (p0 < n) {
            const int nb = n - p0;                            // valid bases in this word (>= 1)
            if (asc[c]) {
                uint32_t x[8];
                if (staged[c]) load32s(stage[wib][c], soff[c] + (uint32_t)p0, x);
                else load32u(in.tig_src + in.tig_off[c][k] + p0, x);
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] &= low_bytes(nb - 4 * i);
                uint32_t wd = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) wd |= not_acgt(x[i]) & low_bytes(nb - 4 * i);
                if (c) weird1 |= wd; else weird0 |= wd;
                planes32(x, H, L);
            } else {
                // one DnaString word = these 32 bases; the half warp's loads are consecutive 8-byte words (in zero-copy
                // mode this is the PCIe traffic of the V..J bases: 1/4 of the ASCII bytes, needed entries only)
                const unsigned long long xw = __ldg(in.tig2_src + in.tig_off[c][k] + wl);
                if (in.tig2_planes) { const uint32_t hi = (uint32_t)(xw >> 32), lo = (uint32_t)xw; H = __brev(hi); L = __brev(hi ^ lo); }
                else planes_of_2bit(xw, H, L);
            }
            const uint32_t valid = low_bits(nb);
            H &= valid; L &= valid;
            const int vend = m.vend[c], jstart = m.jstart[c];
            const uint32_t inV = low_bits(vend - p0) & valid, inJ = valid & ~low_bits(jstart - p0);
            // mutation mask from the germline's own bit planes (k_pack_ref): an ACGT base differs from the germline
            // byte iff the 2-bit codes differ or the germline byte is not ACGT (entries with non-ACGT bases of their
            // own are flagged for the bytewise path and never read their M plane)
            const uint32_t TW = in.ref_plane_words;
            if (inV) {                                        // germline position = p: same word, same alignment
                const uint32_t o = in.ref_plane_off[m.vs[c]] + (uint32_t)wl;
                M |= ((H ^ in.ref_planes[o]) | (L ^ in.ref_planes[TW + o]) | in.ref_planes[2 * TW + o]) & inV;
            }
            if (inJ) {                                        // germline position = p - (n - |j|): any bit offset
                const uint32_t ji = m.js[c];
                const int off = p0 - (n - (int)in.refseq_len[ji]);        // >= -31 for a word that meets the J window
                const int wi = off >> 5;                                  // >= -1: the zero word in front
                const uint32_t bo = (uint32_t)off & 31u;
                const uint32_t *g = in.ref_planes + in.ref_plane_off[ji] + wi;
                const uint32_t gH = __funnelshift_r(g[0], g[1], bo), gL = __funnelshift_r(g[TW], g[TW + 1], bo);
                const uint32_t gN = __funnelshift_r(g[2 * TW], g[2 * TW + 1], bo);
                M |= ((H ^ gH) | (L ^ gL) | gN) & inJ;
            }
            if (c) mcount1 += __popc(M); else mcount0 += __popc(M);
        }
        const int o = (w >> 2) * 12 + (w & 3);
        rows[o] = H; rows[o + 4] = L; rows[o + 8] = M;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) { mcount0 += __shfl_xor_sync(hmask, mcount0, o); mcount1 += __shfl_xor_sync(hmask, mcount1, o); }
    const bool g0 = __any_sync(hmask, weird0 != 0), g1 = __any_sync(hmask, weird1 != 0);
    if (lane == 0) {
        EntryMeta *mm = &meta[s];
        mm->m[0] = (uint16_t)mcount0; mm->m[1] = (uint16_t)mcount1;
        if (g0 || g1) mm->flags = (uint16_t)(mm->flags | (g0 ? FLAG_GENERIC0 : 0u) | (g1 ? FLAG_GENERIC1 : 0u));
    }
}

__global__ void __launch_bounds__(256) k_pack_t2(DevIn in, DevCounters *ctr, EntryMeta *__restrict__ meta, uint4 *__restrict__ t2,
                                                 const uint8_t *__restrict__ needed, unsigned long long wm_lo, unsigned long long wm_hi,
                                                 const uint32_t *__restrict__ asc_list) {
    __shared__ __align__(16) uint8_t stage[16][2][PACK_STAGE];
    const uint32_t hw = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, total = (gridDim.x * blockDim.x) >> 4;
    if (!asc_list) { pack_t2_entry(in, ctr, meta, t2, needed, wm_lo, wm_hi, hw, stage); return; }
    // 2-bit input: only the listed entries (those with an ASCII chain, ~1 %) are this kernel's; a small grid walks the list
    const uint32_t hmask = 0xFFFFu << (threadIdx.x & 16);
    for (uint32_t i = hw; i < ctr->n_asc; i += total) {
        pack_t2_entry(in, ctr, meta, t2, needed, wm_lo, wm_hi, asc_list[i], stage);
        __syncwarp(hmask);                        // the staging buffer is reused by the next entry
    }
}

// The same rows for entries whose two chains both come 2 bits per base (EJOIN_F_TIGS_2BIT; nearly all of them): no bytes to
// convert, so the work is one bit transpose per word and the kernel is organised around its memory traffic instead.
// EIGHT lanes per entry, one 128-base block per lane: four consecutive 8-byte words in (32 contiguous bytes per lane, the
// in-place PCIe traffic in zero-copy mode), three 16-byte rows [H x4 | L x4 | M x4] out (48 contiguous bytes per lane, 16-byte
// stores).  Entries with an ASCII chain are left to k_pack_t2.  Byte tallies: one atomic per warp, no barrier.
__global__ void __launch_bounds__(256) k_pack_t2_2bit(DevIn in, DevCounters *ctr, EntryMeta *__restrict__ meta, uint4 *__restrict__ t2,
                                                      const uint8_t *__restrict__ needed) {
    const uint32_t lane = threadIdx.x & 7;
    unsigned long long my_bytes = 0, my_in = 0;          // tallies: lane 0 of each entry, summed per warp at the end (no block barrier:
                                                         // a barrier made every warp of the CTA wait for its slowest entry)
    const uint32_t gmask = 0xFFu << (threadIdx.x & 24);          // this entry's eight lanes inside the warp
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    bool live = s < (uint32_t)ctr->n_valid && needed[s];
    uint32_t k = 0;
    if (live) { k = meta[s].k; live = in.has_del[0][k] == 0 && in.has_del[1][k] == 0; }
    if (live) {
        const EntryMeta &m = meta[s];
        const int n0 = m.len[0], n1 = m.len[1];
        const uint32_t nb0 = ((uint32_t)n0 + 127) / 128, nblk = nb0 + ((uint32_t)n1 + 127) / 128;
        const unsigned long long off0 = in.tig_off[0][k], off1 = in.tig_off[1][k];
        bool ok = (n0 == 0 || (off0 >= in.tig2_lo && off0 + (unsigned)(n0 + 31) / 32 <= in.tig2_hi)) &&
                  (n1 == 0 || (off1 >= in.tig2_lo && off1 + (unsigned)(n1 + 31) / 32 <= in.tig2_hi));
        if (!ok) { if (lane == 0) atomicOr(&ctr->err, ERR_ARENA_RANGE); }
        else {
            const uint32_t TW = in.ref_plane_words;
            int mcount0 = 0, mcount1 = 0;
            for (uint32_t b = lane; b < nblk; b += 8) {
                const int c = b >= nb0;
                const int wl0 = 4 * (int)(b - (c ? nb0 : 0u));          // first word of this block inside the chain
                const int n = c ? n1 : n0, nwords = (n + 31) / 32;
                const unsigned long long *src = (const unsigned long long *)in.tig2_src + (c ? off1 : off0) + wl0;
                unsigned long long x[4];
#pragma unroll
                for (int q = 0; q < 4; q++) x[q] = wl0 + q < nwords ? __ldg(src + q) : 0ull;      // four loads in flight
                const int vend = m.vend[c], jstart = m.jstart[c];
                const uint32_t vo = in.ref_plane_off[m.vs[c]], ji = m.js[c];
                const int joff0 = 32 * wl0 - (n - (int)in.refseq_len[ji]);
                const uint32_t *jg = in.ref_planes + in.ref_plane_off[ji];
                uint32_t H[4], L[4], M[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int p0 = 32 * (wl0 + q), nb = n - p0;
                    if (in.tig2_planes) { const uint32_t hi = (uint32_t)(x[q] >> 32), lo = (uint32_t)x[q]; H[q] = __brev(hi); L[q] = __brev(hi ^ lo); }
                    else planes_of_2bit(x[q], H[q], L[q]);
                    const uint32_t valid = low_bits(nb);
                    H[q] &= valid; L[q] &= valid;
                    uint32_t mm = 0;
                    const uint32_t inV = low_bits(vend - p0) & valid, inJ = valid & ~low_bits(jstart - p0);
                    if (inV) {
                        const uint32_t o = vo + (uint32_t)(wl0 + q);
                        mm |= ((H[q] ^ in.ref_planes[o]) | (L[q] ^ in.ref_planes[TW + o]) | in.ref_planes[2 * TW + o]) & inV;
                    }
                    if (inJ) {
                        const int off = joff0 + 32 * q;               // >= -31 for a word that meets the J window
                        const int wi = off >> 5;
                        const uint32_t bo = (uint32_t)off & 31u;
                        const uint32_t *g = jg + wi;
                        const uint32_t gH = __funnelshift_r(g[0], g[1], bo), gL = __funnelshift_r(g[TW], g[TW + 1], bo);
                        const uint32_t gN = __funnelshift_r(g[2 * TW], g[2 * TW + 1], bo);
                        mm |= ((H[q] ^ gH) | (L[q] ^ gL) | gN) & inJ;
                    }
                    M[q] = mm;
                    if (c) mcount1 += __popc(mm); else mcount0 += __popc(mm);
                }
                uint4 *rows = t2 + m.t2off + 3 * b;
                rows[0] = make_uint4(H[0], H[1], H[2], H[3]);
                rows[1] = make_uint4(L[0], L[1], L[2], L[3]);
                rows[2] = make_uint4(M[0], M[1], M[2], M[3]);
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) { mcount0 += __shfl_xor_sync(gmask, mcount0, o); mcount1 += __shfl_xor_sync(gmask, mcount1, o); }
            if (lane == 0) {
                EntryMeta *mm = &meta[s];
                mm->m[0] = (uint16_t)mcount0; mm->m[1] = (uint16_t)mcount1;
                my_in = 8ull * ((n0 + 31) / 32) + 8ull * ((n1 + 31) / 32);
                my_bytes = my_in + 48ull * nblk;
            }
        }
    }
    const uint32_t act = __activemask();
    my_bytes = __reduce_add_sync(act, (unsigned)my_bytes);      // < 2^32 per warp: four entries of a few KB each
    my_in = __reduce_add_sync(act, (unsigned)my_in);
    if ((threadIdx.x & 31) == (__ffs(act) - 1) && my_bytes) {
        const uint32_t slot = ((blockIdx.x << 3) | (threadIdx.x >> 5)) & 31;
        atomicAdd(&ctr->pack_bytes[slot], my_bytes);
        atomicAdd(&ctr->inplace_bytes[slot], my_in);
    }
}

// ---------------------------------------------------------------------------------------------
// Decision logic shared by the bit-parallel and the bytewise tier-2 paths (App. A.4 j-n).
// ---------------------------------------------------------------------------------------------
struct PairCounts {
    int cd0, cd1, diffs;
    int shares[2], indeps[2], indeps_out[2];
    int det[2][4];
    int nrefs;
    int dfw, lfw, dc12, l12;     // FWR1 / CDR1+CDR2 region counts; lfw == 0 when the gate is skipped
    int refdiff;                 // positions at which the two entries' assigned references differ (nrefs == 2; help1.rs:345-348)
};

// The decision in two steps, so that callers can count the FWR1 / CDR1+CDR2 differences only for pairs that got past
// the score: decide_score -> accept_reason (0 = reject) with k, d, p1, mult, score; decide_gates -> the two vetoes
// that follow it (A.4 m, n).  `p1tab_L` = tb.p1tab[len0+len1], `mi2` = tb.multinfo[cl0<<8|cl1]: the callers fetch them
// early, together with the meta records.
__device__ __forceinline__ int decide_score(const DevParams &pr, const DevTables &tb, const EntryMeta &m1, const EntryMeta &m2,
                                            const PairCounts &pc, DevCounters *ctr, const double *p1tab_L, const MultInfo mi2, int *kout,
                                            int *dout, int *mo_out, double *p1o, double *multo, double *scoreo) {
    const int cd = pc.cd0 + pc.cd1;
    // The host no longer waits for the key pass to learn which tables an input needs (one sync per join): a pair whose
    // table is not there yet raises a flag, the pipeline's results are dropped, the tables are built and the join reruns.
    if (p1tab_L == nullptr || mi2.off == 0xFFFFFFFFu) { atomicOr(&ctr->err, ERR_TABLES_MISSING); return 0; }
    if (pc.nrefs == 2) {            // MAX_DEGRADATION, both readings (include/enclone_join.h: degradation_mode)
        if (pr.degradation_mode == 1) { int df = pc.shares[0] - pc.shares[1]; if (df < 0) df = -df; if (df > pr.max_degradation) return 0; }
        else if (pc.refdiff > pr.max_degradation) return 0;
    }
    int d = pc.shares[0], mi = pc.indeps[0], mo = pc.indeps_out[0];
    if (pc.nrefs == 2) { d = min(d, pc.shares[1]); mi = min(mi, pc.indeps[1]); mo = min(mo, pc.indeps_out[1]); }
    const int k = mi + 2 * d;
    double p1 = 1.0;
    if (d >= P1_DCAP) p1 = 0.0;          // below the reference's double-double noise floor
    else if (d > 0) {
        if (k >= P1_KCAP) atomicOr(&ctr->err, ERR_K_TOO_LARGE);
        else p1 = p1tab_L[d * P1_KCAP + k];
    }
    const double mult = tb.multpool[mi2.off + (uint32_t)pc.cd0 * mi2.d1p1 + (uint32_t)pc.cd1];
    const double score = __dmul_rn(p1, mult);
    int reason = 0;
    if (!(score > pr.max_score)) reason |= 1;
    if (d >= pr.auto_share) reason |= 2;
    {
        int hc = min((int)m1.hcomp, (int)m2.hcomp);
        if (hc - cd >= pr.comp_filt && pc.diffs - cd <= pr.comp_filt_bound) reason |= 4;
    }
    *kout = k; *dout = d; *mo_out = mo; *p1o = p1; *multo = mult; *scoreo = score;
    return reason;
}
__device__ __forceinline__ bool decide_gates(const DevParams &pr, const PairCounts &pc, int mo) {
    if (pr.is_bcr && pc.lfw > 0) {
        const double ifw = __dmul_rn(100.0, __dsub_rn(1.0, __ddiv_rn((double)pc.dfw, (double)pc.lfw)));
        const double i12 = __dmul_rn(100.0, __dsub_rn(1.0, __ddiv_rn((double)pc.dc12, (double)pc.l12)));
        if (__dsub_rn(ifw, i12) >= pr.fwr1_cdr12_delta) return false;
    }
    const int x = mo > 1 ? mo : 1;
    return !(pc.cd0 + pc.cd1 >= pr.cdr3_mult * x);
}
// returns accept_reason (0 = reject); fills k, d, p1, mult, score
__device__ __forceinline__ int decide(const DevParams &pr, const DevTables &tb, const EntryMeta &m1, const EntryMeta &m2,
                                      const PairCounts &pc, DevCounters *ctr, int *kout, int *dout, double *p1o, double *multo,
                                      double *scoreo) {
    int mo = 0;
    const int reason = decide_score(pr, tb, m1, m2, pc, ctr, tb.p1tab[(int)m1.len[0] + (int)m1.len[1]],
                                    tb.multinfo[((uint32_t)m1.cl[0] << 8) | m1.cl[1]], kout, dout, &mo, p1o, multo, scoreo);
    if (!reason) return 0;
    return decide_gates(pr, pc, mo) ? reason : 0;
}

// cheap gates that need only the meta records (App. A.4 e-g); returns 0 = reject, 1 = pass; *err = cross-donor
__device__ __forceinline__ int meta_gates(const DevParams &pr, const DevIn &in, const EntryMeta &m1, const EntryMeta &m2, int cd, int *err) {
    int e = 0;
    if (m1.donor_lo != EJOIN_NONE && m2.donor_lo != EJOIN_NONE)
        e = !(m1.donor_lo == m1.donor_hi && m2.donor_lo == m2.donor_hi && m1.donor_lo == m2.donor_lo);
    *err = e;
    if (e && !pr.mix_donors) return 0;
    if (pr.is_bcr && !pr.old_light && cd > 0 && m1.c_light != m2.c_light) return 0;
    for (int c = 0; c < 2; c++) {
        if (m1.vname[c] == m2.vname[c]) continue;
        // V names differ: universal V references must agree after right-truncation, and the
        // 5'UTR references (when both known) after left-truncation (help1.rs:328-331)
        uint32_t a = in.vuni_seq[c][m1.k], b = in.vuni_seq[c][m2.k];
        if (a >= in.n_refseq || b >= in.n_refseq) return 0;
        uint32_t la = in.refseq_len[a], lb = in.refseq_len[b], mlen = la < lb ? la : lb;
        const uint8_t *ra = in.refseq_arena + in.refseq_off[a], *rb = in.refseq_arena + in.refseq_off[b];
        for (uint32_t i = 0; i < mlen; i++) if (ra[i] != rb[i]) return 0;
        uint32_t ua = in.utr_seq[c][m1.k], ub = in.utr_seq[c][m2.k];
        if (ua != EJOIN_NONE && ub != EJOIN_NONE && ua < in.n_refseq && ub < in.n_refseq) {
            uint32_t lua = in.refseq_len[ua], lub = in.refseq_len[ub], mu = lua < lub ? lua : lub;
            const uint8_t *pa = in.refseq_arena + in.refseq_off[ua] + (lua - mu), *pb = in.refseq_arena + in.refseq_off[ub] + (lub - mu);
            for (uint32_t i = 0; i < mu; i++) if (pa[i] != pb[i]) return 0;
        }
    }
    return 1;
}

__device__ __forceinline__ bool region_bounds(const EntryMeta &a, const EntryMeta &b, int *f1, int *c1, int *f2, int *c2, int *f3) {
    // k1's bounds, provided both entries know all of theirs (oracle_fwr1_cdr12_counts)
    if (a.fr1 == EJOIN_NONE16 || a.cdr1 == EJOIN_NONE16 || a.fr2 == EJOIN_NONE16 || a.cdr2 == EJOIN_NONE16 || a.fr3 == EJOIN_NONE16) return false;
    if (b.fr1 == EJOIN_NONE16 || b.cdr1 == EJOIN_NONE16 || b.fr2 == EJOIN_NONE16 || b.cdr2 == EJOIN_NONE16 || b.fr3 == EJOIN_NONE16) return false;
    *f1 = a.fr1; *c1 = a.cdr1; *f2 = a.fr2; *c2 = a.cdr2; *f3 = a.fr3;
    if (!(*f1 < *c1 && *c1 <= *f2 && *f2 <= *c2 && *c2 <= *f3 && *f3 <= (int)a.len[0] && (*f2 - *c1) + (*f3 - *c2) > 0)) return false;
    return true;
}

__device__ __forceinline__ void emit_raw(const DevParams &pr, DevCounters *ctr, ejoin_raw_edge_t *acc, unsigned long long cap,
                                         uint32_t k1, uint32_t k2, int cd, int d, int reason, int err, bool accepted) {
    // warp-aggregated append
    const uint32_t lane = threadIdx.x & 31;
    uint32_t mask = __ballot_sync(__activemask(), accepted);
    if (!accepted) return;
    uint32_t leader = __ffs(mask) - 1, rank = __popc(mask & ((1u << lane) - 1u));
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&ctr->n_accept, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader) + rank;
    if (base < cap) {
        ejoin_raw_edge_t e;
        e.k1 = k1; e.k2 = k2; e.cd = (uint16_t)cd; e.d = (uint16_t)min(d, 65535);
        e.flags = (uint32_t)reason | ((uint32_t)err << 8);
        acc[base] = e;
    }
}

// ---------------------------------------------------------------------------------------------
// Thread-level full evaluation of one pair (positions sa, sb of one comparison group; the entry
// with the smaller info index is k1).  Bit-parallel over the t2 rows.  Returns accept_reason
// (0 = reject) or -1 when the pair needs the bytewise path (has_del / non-ACGT / overlapping
// windows / different germlines); *scored is set when the pair got past the cheap gates.
// ---------------------------------------------------------------------------------------------
// CDR3 differences of a pair from its two c3 rows, all 4W words in flight at once
template <int W>
__device__ __forceinline__ void cdr3_counts_w(const uint32_t *__restrict__ c3, uint32_t o1, uint32_t o2, int cl0, int *cd0, int *cdt) {
    uint32_t a[2 * W], b[2 * W];
#pragma unroll
    for (int i = 0; i < 2 * W; i++) { a[i] = __ldg(c3 + o1 + i); b[i] = __ldg(c3 + o2 + i); }
    int t = 0, t0 = 0;
#pragma unroll
    for (int w = 0; w < W; w++) {
        const uint32_t D = (a[w] ^ b[w]) | (a[W + w] ^ b[W + w]);
        t += __popc(D);
        t0 += __popc(D & range_mask(w, 0, cl0));
    }
    *cd0 = t0; *cdt = t;
}
union MetaRegs { EntryMeta m; uint4 q[8]; __device__ MetaRegs() {} };

__device__ __forceinline__ int evaluate_thread(const DevIn &in, const DevTables &tb, const DevParams &pr, DevCounters *ctr, uint32_t sa,
                                               uint32_t sb, int *cd_out, int *d_out, int *err_out, bool *scored, unsigned long long *bytes) {
    // Inside a comparison group sorted position order == info order (stable sort of ascending k), so the
    // entry at the smaller position is k1: no need to look at the records to order the pair.
    if (sa > sb) { const uint32_t t = sa; sa = sb; sb = t; }
    MetaRegs u1, u2;                      // both 128-byte records with 16 loads in flight, then everything from registers
    {
        const uint4 *q1 = (const uint4 *)&tb.meta[sa], *q2 = (const uint4 *)&tb.meta[sb];
#pragma unroll
        for (int i = 0; i < 8; i++) { u1.q[i] = __ldg(q1 + i); u2.q[i] = __ldg(q2 + i); }
    }
    const EntryMeta &m1 = u1.m, &m2 = u2.m;
    *scored = false; *d_out = 0; *err_out = 0;
    // table handles of this pair's (n, CDR3 lengths): fetched now, used by the score at the very end
    const double *p1tab_L = tb.p1tab[(int)m1.len[0] + (int)m1.len[1]];
    const MultInfo mi2 = tb.multinfo[((uint32_t)m1.cl[0] << 8) | m1.cl[1]];
    // CDR3 differences per chain from the c3 rows
    const uint32_t cl0 = m1.cl[0], total = cl0 + m1.cl[1], W = max(1u, (total + 31) / 32);
    int cd0 = 0, cdt = 0;
    switch (W) {
        case 1: cdr3_counts_w<1>(tb.c3, m1.c3off, m2.c3off, (int)cl0, &cd0, &cdt); break;
        case 2: cdr3_counts_w<2>(tb.c3, m1.c3off, m2.c3off, (int)cl0, &cd0, &cdt); break;
        case 3: cdr3_counts_w<3>(tb.c3, m1.c3off, m2.c3off, (int)cl0, &cd0, &cdt); break;
        case 4: cdr3_counts_w<4>(tb.c3, m1.c3off, m2.c3off, (int)cl0, &cd0, &cdt); break;
        default:
            for (uint32_t w = 0; w < W; w++) {
                uint32_t D = (tb.c3[m1.c3off + w] ^ tb.c3[m2.c3off + w]) | (tb.c3[m1.c3off + W + w] ^ tb.c3[m2.c3off + W + w]);
                cdt += __popc(D);
                cd0 += __popc(D & range_mask((int)w, 0, (int)cl0));
            }
    }
    const int cd = cdt, cd1 = cdt - cd0;
    *cd_out = cd;
    const uint32_t cdm = tb.cdmax[total];
    if (cdm == 0xFFFF || (uint32_t)cd > cdm) return 0;          // CDR3 gates (A.4-d), same table as tier 1
    if (!meta_gates(pr, in, m1, m2, cd, err_out)) return 0;
    *scored = true;
    const uint32_t nb0 = ((uint32_t)m1.len[0] + 127) / 128, nblk = nb0 + ((uint32_t)m1.len[1] + 127) / 128;
    const bool generic = ((m1.flags | m2.flags) & (FLAG_GENERIC0 | FLAG_GENERIC1)) || m1.vs[0] != m2.vs[0] || m1.vs[1] != m2.vs[1] ||
                         m1.js[0] != m2.js[0] || m1.js[1] != m2.js[1];
    if (generic) { *bytes += 2ull * (48ull * nblk + 128ull); return -1; }
    {
        // No shared mutation is possible when one of the entries has none (shares <= each entry's own count of germline
        // differences inside the windows): then d = 0, p1 = 1 and the score is the CDR3 multiplier alone.  If that
        // exceeds MAX_SCORE, AUTO_SHARES cannot fire (d = 0) and the junction-complexity exception is out of reach
        // (hcomp - cd < COMP_FILT), join_one rejects the pair whatever the rest of the sequences say: no t2 row is read.
        // (Naive cells with similar CDR3s: most of the rejected pairs of a repertoire.)
        const int mt1 = (int)m1.m[0] + (int)m1.m[1], mt2 = (int)m2.m[0] + (int)m2.m[1];
        if (min(mt1, mt2) == 0 && pr.auto_share > 0 && min((int)m1.hcomp, (int)m2.hcomp) - cd < pr.comp_filt) {
            const double mult0 = tb.multpool[mi2.off + (uint32_t)cd0 * mi2.d1p1 + (uint32_t)cd1];
            if (__dmul_rn(1.0, mult0) > pr.max_score) { *bytes += 256ull; return 0; }
        }
    }
    *bytes += 2ull * (48ull * nblk + 128ull);
    PairCounts pc;
    pc.cd0 = cd0; pc.cd1 = cd1; pc.nrefs = 1; pc.refdiff = 0;      // same germline sequences on this path
    for (int c2 = 0; c2 < 2; c2++)
        if (m1.vref[c2] != m2.vref[c2] || m1.dref[c2] != m2.dref[c2] || m1.jref[c2] != m2.jref[c2]) pc.nrefs = 2;
    // One pass over the t2 rows counts total differences, shared mutations and, when cd >= CDR3_MULT, the non-shared
    // mutations outside the CDR3s (A.4-n).  The FWR1 vs CDR1+CDR2 differences of the heavy chain (A.4-m) only veto
    // pairs that got past the score (one in seven in phase B), so they are counted afterwards, for those pairs, from
    // the few blocks that meet [f1,f3) — by then in L1.
    const bool do_io = cd >= pr.cdr3_mult;
    int diffs = 0, shares = 0, io = 0;
    const uint4 *r1 = tb.t2 + m1.t2off, *r2 = tb.t2 + m2.t2off;
#pragma unroll 2
    for (uint32_t b = 0; b < nblk; b++) {
        const uint4 h1 = __ldg(r1 + 3 * b), l1 = __ldg(r1 + 3 * b + 1), x1 = __ldg(r1 + 3 * b + 2);
        const uint4 h2 = __ldg(r2 + 3 * b), l2 = __ldg(r2 + 3 * b + 1), x2 = __ldg(r2 + 3 * b + 2);
        const uint32_t D[4] = { (h1.x ^ h2.x) | (l1.x ^ l2.x), (h1.y ^ h2.y) | (l1.y ^ l2.y), (h1.z ^ h2.z) | (l1.z ^ l2.z), (h1.w ^ h2.w) | (l1.w ^ l2.w) };
        const uint32_t X1[4] = { x1.x, x1.y, x1.z, x1.w }, X2[4] = { x2.x, x2.y, x2.z, x2.w };
#pragma unroll
        for (int q = 0; q < 4; q++) { diffs += __popc(D[q]); shares += __popc(X1[q] & X2[q] & ~D[q]); }
        const bool ch1 = b >= nb0;
        const int wl0 = 4 * (int)(b - (ch1 ? nb0 : 0u));          // word index inside the chain
        if (do_io) {
            const int a1 = ch1 ? m1.cs[1] : m1.cs[0], a2 = ch1 ? m2.cs[1] : m2.cs[0], cl = ch1 ? m1.cl[1] : m1.cl[0];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t R = range_mask(wl0 + q, a1, a1 + cl) | range_mask(wl0 + q, a2, a2 + cl);
                io += __popc(X1[q] & D[q] & ~R) + __popc(X2[q] & D[q] & ~R);
            }
        }
    }
    const int mm = (int)m1.m[0] + (int)m1.m[1] + (int)m2.m[0] + (int)m2.m[1];
    pc.diffs = diffs;
    pc.shares[0] = pc.shares[1] = shares;
    pc.indeps[0] = pc.indeps[1] = mm - 2 * shares;
    pc.indeps_out[0] = pc.indeps_out[1] = do_io ? io : (1 << 20);    // only looked at when the CDR3_MULT rule can bite
    // total-difference cap (A.4-c)
    if (!pr.is_bcr || pr.max_diffs < 1000000) {
        int cap = pr.max_diffs;
        if (!pr.is_bcr && pr.max_diffs_tcr < cap) cap = pr.max_diffs_tcr;
        if (diffs > cap) return 0;
    }
    int kk, mo; double p1, mult, score;
    const int reason = decide_score(pr, tb, m1, m2, pc, ctr, p1tab_L, mi2, &kk, d_out, &mo, &p1, &mult, &score);
    if (!reason) return 0;
    pc.lfw = 0; pc.dfw = pc.dc12 = pc.l12 = 0;
    int f1 = 0, c1 = 0, f2 = 0, c2_ = 0, f3 = 0;
    if (pr.is_bcr && region_bounds(m1, m2, &f1, &c1, &f2, &c2_, &f3)) {
        int dfw = 0, d12 = 0;
        for (int b = f1 >> 7; b <= (f3 - 1) >> 7; b++) {             // f3 <= len0: blocks of chain 0
            const uint4 h1 = __ldg(r1 + 3 * b), l1 = __ldg(r1 + 3 * b + 1), h2 = __ldg(r2 + 3 * b), l2 = __ldg(r2 + 3 * b + 1);
            const uint32_t D[4] = { (h1.x ^ h2.x) | (l1.x ^ l2.x), (h1.y ^ h2.y) | (l1.y ^ l2.y), (h1.z ^ h2.z) | (l1.z ^ l2.z), (h1.w ^ h2.w) | (l1.w ^ l2.w) };
#pragma unroll
            for (int q = 0; q < 4; q++) {
                dfw += __popc(D[q] & range_mask(4 * b + q, f1, c1));
                d12 += __popc(D[q] & (range_mask(4 * b + q, c1, f2) | range_mask(4 * b + q, c2_, f3)));
            }
        }
        pc.dfw = dfw; pc.lfw = c1 - f1; pc.dc12 = d12; pc.l12 = (f2 - c1) + (f3 - c2_);
    }
    return decide_gates(pr, pc, mo) ? reason : 0;
}

// ---------------------------------------------------------------------------------------------
// K5  tier 2, phase B.  k_phaseb_split (pass 1 over the recorded survivors): the survivors of entries
// that gave up in phase A (force) go to `todo` (all scored in round 1); the others are kept, compacted,
// in `rest` unless they visibly hang off the same phase-A parent (same tree: never a forest edge).
// k_phaseb_cross (pass 2, after round 1 and the pointer doubling) keeps from `rest` the pairs whose
// endpoints lie in different trees.  k_tier2: one thread per kept pair, full evaluation.
// ---------------------------------------------------------------------------------------------
// Block-level compaction for the list filters below: every thread brings up to SPLIT_ITEMS flagged items for each of
// two output lists; one atomicAdd per list and block tile reserves the slots (a per-warp atomic on one counter was the
// whole cost of these kernels: ~1M same-address atomics per launch).
#define SPLIT_ITEMS 4
struct SplitSmem { uint32_t w1[8], w2[8]; unsigned long long base1, base2; };
__device__ __forceinline__ void block_reserve2(SplitSmem &sm, uint32_t c1, uint32_t c2, unsigned long long *ctr1, unsigned long long *ctr2,
                                               unsigned long long &o1, unsigned long long &o2) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t i1 = c1, i2 = c2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v1 = __shfl_up_sync(0xFFFFFFFFu, i1, o), v2 = __shfl_up_sync(0xFFFFFFFFu, i2, o);
        if (lane >= (uint32_t)o) { i1 += v1; i2 += v2; }
    }
    if (lane == 31) { sm.w1[warp] = i1; sm.w2[warp] = i2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t1 = 0, t2 = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const uint32_t a = sm.w1[w], b = sm.w2[w]; sm.w1[w] = t1; sm.w2[w] = t2; t1 += a; t2 += b; }
        sm.base1 = t1 ? atomicAdd(ctr1, (unsigned long long)t1) : 0ull;
        sm.base2 = (t2 && ctr2) ? atomicAdd(ctr2, (unsigned long long)t2) : 0ull;
    }
    __syncthreads();
    o1 = sm.base1 + sm.w1[warp] + (i1 - c1);
    o2 = sm.base2 + sm.w2[warp] + (i2 - c2);
    __syncthreads();                                  // sm is reused by the next tile
}

__global__ void __launch_bounds__(256) k_phaseb_split(const uint4 *__restrict__ cand, unsigned long long cand_cap, const uint32_t *__restrict__ parent,
                                                      const uint8_t *__restrict__ force_b, DevCounters *ctr, uint2 *__restrict__ todo,
                                                      unsigned long long todo_cap, uint2 *__restrict__ rest, unsigned long long rest_cap) {
    __shared__ SplitSmem sm;
    unsigned long long ncand = ctr->n_cand;
    if (ncand > cand_cap) ncand = cand_cap;
    // one record (b, 32 a's as a mask) per thread and tile: the mask splits into the pairs scored in round 1 (b gave up in
    // phase A) and the pairs kept for round 2 (endpoints not visibly in one phase-A tree)
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * 256ull; t0 < ncand; t0 += (unsigned long long)gridDim.x * 256ull) {
        const unsigned long long i = t0 + threadIdx.x;
        uint4 c = make_uint4(0, 0, 0, 0);
        uint32_t k1 = 0, k2 = 0;                      // masks: bit j -> pair (c.x + j, c.y) goes to todo / rest
        if (i < ncand) {
            c = cand[i];
            if ((c.w & 1u) || force_b[c.y]) k1 = c.z;
            else {
                const uint32_t pb = parent[c.y];
                uint32_t m = c.z;
                while (m) {
                    const uint32_t j = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t a = c.x + j;
                    if (parent[a] != pb && pb != a) k2 |= 1u << j;
                }
            }
        }
        unsigned long long o1, o2;
        block_reserve2(sm, __popc(k1), __popc(k2), &ctr->n_todo, &ctr->n_rest, o1, o2);
        while (k1) { const uint32_t j = __ffs(k1) - 1; k1 &= k1 - 1; if (o1 < todo_cap) todo[o1] = make_uint2(c.x + j, c.y); o1++; }
        while (k2) { const uint32_t j = __ffs(k2) - 1; k2 &= k2 - 1; if (o2 < rest_cap) rest[o2] = make_uint2(c.x + j, c.y); o2++; }
    }
}
__global__ void __launch_bounds__(256) k_phaseb_cross(const uint2 *__restrict__ rest, unsigned long long rest_cap, const uint32_t *__restrict__ root,
                                                      DevCounters *ctr, uint2 *__restrict__ todo, unsigned long long todo_cap) {
    __shared__ SplitSmem sm;
    unsigned long long n = ctr->n_rest;
    if (n > rest_cap) n = rest_cap;
    const unsigned long long tile = 256ull * SPLIT_ITEMS;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * tile; t0 < n; t0 += (unsigned long long)gridDim.x * tile) {
        uint2 c[SPLIT_ITEMS];
        uint32_t k1 = 0;
#pragma unroll
        for (int j = 0; j < SPLIT_ITEMS; j++) {
            const unsigned long long i = t0 + (unsigned long long)j * 256 + threadIdx.x;
            c[j] = make_uint2(0, 0);
            if (i < n) { c[j] = rest[i]; if (root[c[j].x] != root[c[j].y]) k1 |= 1u << j; }
        }
        unsigned long long o1, o2;
        block_reserve2(sm, __popc(k1), 0, &ctr->n_todo, nullptr, o1, o2);
#pragma unroll
        for (int j = 0; j < SPLIT_ITEMS; j++)
            if ((k1 >> j) & 1u) { if (o1 < todo_cap) todo[o1] = c[j]; o1++; }
    }
}

__global__ void __launch_bounds__(128, 5) k_tier2(DevIn in, DevTables tb, DevParams pr, DevCounters *ctr, const uint2 *todo,
                                               unsigned long long cand_cap, uint2 *generic_q, ejoin_raw_edge_t *acc, uint32_t k_base,
                                               uint32_t *parent_min) {
    unsigned long long ntodo = ctr->n_todo;
    if (ntodo > cand_cap) ntodo = cand_cap;
    unsigned long long my_pairs = 0, my_bytes = 0;      // per-thread tallies, one atomic per warp at the end
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < ((ntodo + 31) & ~31ull);
         i += (unsigned long long)gridDim.x * blockDim.x) {
        bool accepted = false;
        uint32_t k1 = 0, k2 = 0;
        int cd = 0, d = 0, reason = 0, err = 0;
        if (i < ntodo) {
            const uint2 c = todo[i];
            bool scored;
            reason = evaluate_thread(in, tb, pr, ctr, c.x, c.y, &cd, &d, &err, &scored, &my_bytes);
            my_pairs += scored;
            if (reason < 0) {
                unsigned long long q = atomicAdd(&ctr->n_generic_q, 1ull);
                if (q < cand_cap) generic_q[q] = c;            // c.x < c.y positions: already k1 < k2
                reason = 0;
            }
            accepted = reason != 0;
            k1 = tb.meta[c.x].k; k2 = tb.meta[c.y].k;
            if (k1 > k2) { uint32_t t = k1; k1 = k2; k2 = t; }
            // round 1: the smallest accepted partner of b becomes its phase-A parent after all
            if (accepted && parent_min) atomicMin(&parent_min[c.y], c.x);
        }
        emit_raw(pr, ctr, acc, cand_cap, k1 + k_base, k2 + k_base, cd, d, reason, err, accepted);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_pairs += __shfl_xor_sync(0xFFFFFFFFu, my_pairs, o);
        my_bytes += __shfl_xor_sync(0xFFFFFFFFu, my_bytes, o);
    }
    if ((threadIdx.x & 31) == 0 && my_pairs) { atomicAdd(&ctr->pairs_tier2, my_pairs); atomicAdd(&ctr->t2_alg_bytes, my_bytes); }
}

// After round 1: an accepted edge (a,b) of a gave-up b is b's tree edge iff a is the minimum
// (parent[b] == a): move it to the phase-A slot.  Runs before the pointer doubling.
__global__ void k_forced_resolve(ejoin_raw_edge_t *acc, DevCounters *ctr, unsigned long long cap, const uint32_t *k_to_s, uint32_t k_base,
                                 const uint32_t *parent, unsigned long long *f0_key, unsigned long long *f0_val) {
    unsigned long long n = ctr->n_accept;
    if (n > cap) n = cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        ejoin_raw_edge_t e = acc[i];
        const uint32_t sa = k_to_s[e.k1 - k_base], sb = k_to_s[e.k2 - k_base];
        if (parent[sb] == sa) {
            f0_key[sb] = ((unsigned long long)e.k1 << 32) | e.k2;
            f0_val[sb] = ((unsigned long long)e.flags << 32) | ((unsigned long long)e.d << 16) | e.cd;
            acc[i].k1 = 0xFFFFFFFFu;                       // dead: now a tree edge
        }
    }
}
// Between the rounds: remember where round 1 ended and empty the todo list.
__global__ void k_round2_begin(DevCounters *ctr) {
    ctr->n_accept_r1 = ctr->n_accept;
    ctr->n_generic_r1 = ctr->n_generic_q;
    ctr->n_todo_r1 = ctr->n_todo;
    ctr->n_todo = 0;
}
// After the pointer doubling: the remaining round-1 edges inside one tree are the heaviest edge of
// a cycle and are dropped.
__global__ void k_forced_prune(ejoin_raw_edge_t *acc, const DevCounters *ctr, unsigned long long cap, const uint32_t *k_to_s,
                               uint32_t k_base, const uint32_t *root) {
    unsigned long long n_round1 = ctr->n_accept_r1;
    if (n_round1 > cap) n_round1 = cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round1; i += (unsigned long long)gridDim.x * blockDim.x) {
        ejoin_raw_edge_t e = acc[i];
        if (e.k1 == 0xFFFFFFFFu) continue;
        if (root[k_to_s[e.k1 - k_base]] == root[k_to_s[e.k2 - k_base]]) acc[i].k1 = 0xFFFFFFFFu;
    }
}

// ---------------------------------------------------------------------------------------------
// Bytewise evaluator: one warp per pair, straight from the ASCII bytes — for entries the
// bit-parallel path cannot represent (has_del / non-ACGT / overlapping windows), for pairs
// whose germlines differ (nrefs = 2 with explicit references), and for the full payload of
// the emitted edges.  Mirrors join_one's loops (App. A.4 c,d,i,m).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// 16 bytes starting at an arbitrary address, from two aligned 16-byte loads (every arena carries >= 64 bytes of
// slack behind its last byte).  out[i] holds bytes 4i..4i+3.
__device__ __forceinline__ void load16u(const uint8_t *p, uint32_t (&out)[4]) {
    const uintptr_t a = (uintptr_t)p;
    const uint4 *q = (const uint4 *)(a & ~(uintptr_t)15);
    const uint4 v0 = __ldg(q), v1 = __ldg(q + 1);
    const uint32_t w[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
    const uint32_t sh = 8u * (uint32_t)(a & 3);
    switch ((a >> 2) & 3) {
        case 0:
#pragma unroll
            for (int i = 0; i < 4; i++) out[i] = __funnelshift_r(w[i], w[i + 1], sh);
            break;
        case 1:
#pragma unroll
            for (int i = 0; i < 4; i++) out[i] = __funnelshift_r(w[i + 1], w[i + 2], sh);
            break;
        case 2:
#pragma unroll
            for (int i = 0; i < 4; i++) out[i] = __funnelshift_r(w[i + 2], w[i + 3], sh);
            break;
        default:
#pragma unroll
            for (int i = 0; i < 3; i++) out[i] = __funnelshift_r(w[i + 3], w[i + 4], sh);
            out[3] = __funnelshift_r(w[6], w[7], sh);
            break;
    }
}
// bit i = (byte i of a differs from byte i of b), i < 16
__device__ __forceinline__ uint32_t neq16(const uint32_t (&a)[4], const uint32_t (&b)[4]) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) m |= ((nonzero_bytes(a[i] ^ b[i]) * 0x01020408u) >> 24 & 0xFu) << (4 * i);
    return m;
}
// bits i < 16 with lo <= p0 + i < hi
__device__ __forceinline__ uint32_t range16(int p0, int lo, int hi) {
    int a = lo - p0, b = hi - p0;
    a = a < 0 ? 0 : a;
    b = b > 16 ? 16 : b;
    if (b <= a) return 0u;
    return ((1u << b) - 1u) & ~((1u << a) - 1u);
}

// The 32 lanes of a warp take 16 consecutive positions each (one pass covers 512 bases): the bytes of both entries and
// of the germlines are fetched with 16-byte loads, compared as 16-bit masks, and every count of join_one's loops is a
// popcount of mask products.  The first version walked the sequences one byte per lane and pass (diffs, V window, J
// window, per germline set): ~60 dependent load rounds per pair, against 3 here.
template <bool WARP>
__device__ void bytewise_counts_t(const DevIn &in, const uint4 *__restrict__ t2rows, const uint32_t *__restrict__ c3rows, const DevParams &pr, const EntryMeta &m1, const EntryMeta &m2, PairCounts &pc) {
    static_assert(WARP, "warp-cooperative only");
    const int lane = (int)(threadIdx.x & 31);
    const uint32_t k1 = m1.k, k2 = m2.k;
    int nrefs = 1;
    for (int c = 0; c < 2; c++)
        if (m1.vref[c] != m2.vref[c] || m1.dref[c] != m2.dref[c] || m1.jref[c] != m2.jref[c]) nrefs = 2;
    pc.nrefs = nrefs;
    int diffs = 0, cdc[2] = { 0, 0 };
    int sh[2] = { 0, 0 }, in_[2] = { 0, 0 }, io[2] = { 0, 0 }, det[2][4] = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 } };
    int dfw = 0, d12 = 0;
    int f1 = 0, c1 = 0, f2 = 0, c2 = 0, f3 = 0;
    const bool reg = pr.is_bcr && region_bounds(m1, m2, &f1, &c1, &f2, &c2, &f3);
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int n = m1.len[c];
        // the bytes of a 2-bit chain are rebuilt from its plane words in the t2 rows
        const bool asc1 = in.tig2_src == nullptr || in.has_del[c][k1] != 0, asc2 = in.tig2_src == nullptr || in.has_del[c][k2] != 0;
        const uint8_t *t1 = in.tig_arena + (asc1 ? in.tig_off[c][k1] : 0ull), *t2 = in.tig_arena + (asc2 ? in.tig_off[c][k2] : 0ull);
        const uint32_t blk0 = c ? ((uint32_t)m1.len[0] + 127) / 128 : 0u;
        const uint32_t *w1 = (const uint32_t *)(t2rows + m1.t2off) + 12 * blk0, *w2 = (const uint32_t *)(t2rows + m2.t2off) + 12 * blk0;
        const int a1 = m1.cs[c], a2 = m2.cs[c], cl = m1.cl[c];
        if (c == 0 && lane < 8) {   // CDR3 differences per chain from the c3 rows (the CDR3s are ACGT only): lane w takes word w
            const int cl0 = m1.cl[0], total = cl0 + (int)m1.cl[1], W = max(1, (total + 31) / 32);
            if (lane < W) {
                const uint32_t D = (__ldg(c3rows + m1.c3off + lane) ^ __ldg(c3rows + m2.c3off + lane)) |
                                   (__ldg(c3rows + m1.c3off + W + lane) ^ __ldg(c3rows + m2.c3off + W + lane));
                const int d0 = __popc(D & range_mask(lane, 0, cl0));
                cdc[0] += d0; cdc[1] += __popc(D) - d0;
            }
        }
        for (int p0 = 16 * lane; p0 < n; p0 += 512) {
            uint32_t x[4], y[4];
            {
                const int o = (p0 >> 7) * 12 + ((p0 >> 5) & 3), half = (p0 >> 4) & 1;
                if (asc1) load16u(t1 + p0, x); else bytes_of_planes16(__ldg(w1 + o), __ldg(w1 + o + 4), half, x);
                if (asc2) load16u(t2 + p0, y); else bytes_of_planes16(__ldg(w2 + o), __ldg(w2 + o + 4), half, y);
            }
            const uint32_t valid = range16(p0, 0, n);
            const uint32_t D = neq16(x, y) & valid, S = ~D & valid;          // positions where the entries differ / agree
            const uint32_t out = valid & ~(range16(p0, a1, a1 + cl) | range16(p0, a2, a2 + cl));
            diffs += __popc(D);
            if (reg && c == 0) { dfw += __popc(D & range16(p0, f1, c1)); d12 += __popc(D & (range16(p0, c1, f2) | range16(p0, c2, f3))); }
            for (int u = 0; u < nrefs; u++) {
                const EntryMeta &mu = u == 0 ? m1 : m2;
                const uint32_t vi = mu.vs[c], ji = mu.js[c];
                const int vl = (int)in.refseq_len[vi], jl = (int)in.refseq_len[ji];
                const uint32_t inV = range16(p0, 0, min(vl - pr.ref_v_trim, n));
                if (inV) {
                    uint32_t r[4];
                    load16u(in.refseq_arena + in.refseq_off[vi] + p0, r);
                    const uint32_t A = neq16(x, r) & inV, B = neq16(y, r) & inV;
                    const int s_ = __popc(S & A);
                    sh[u] += s_; det[u][2 * c] += s_;
                    in_[u] += __popc(D & A) + __popc(D & B);
                    io[u] += __popc(D & A & out) + __popc(D & B & out);
                }
                const int jw = jl - pr.ref_j_trim;
                const uint32_t inJ = jw > 0 ? range16(p0, max(0, n - jw), n) : 0u;
                if (inJ) {
                    const int g0 = p0 - (n - jl);                 // germline index of position p0
                    const uint8_t *jp = in.refseq_arena + in.refseq_off[ji];
                    uint32_t r[4];
                    if (g0 >= 0) load16u(jp + g0, r);
                    else {                                        // the chunk starts before the germline does: byte by byte
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            uint32_t wv = 0;
#pragma unroll
                            for (int b = 0; b < 4; b++) { const int g = g0 + 4 * i + b; if (g >= 0 && g < jl) wv |= (uint32_t)jp[g] << (8 * b); }
                            r[i] = wv;
                        }
                    }
                    const uint32_t A = neq16(x, r) & inJ, B = neq16(y, r) & inJ;
                    const int s_ = __popc(S & A);
                    sh[u] += s_; det[u][2 * c + 1] += s_;
                    in_[u] += __popc(D & A) + __popc(D & B);
                    io[u] += __popc(D & A & out) + __popc(D & B & out);
                }
            }
        }
    }
    int refdiff = 0;
    if (nrefs == 2 && pr.degradation_mode != 1) {
        // V references left-aligned, J references right-aligned; unmatched positions of the longer one count as differing
        for (int c = 0; c < 2; c++) {
            for (int vj = 0; vj < 2; vj++) {
                const uint32_t i1 = vj ? m1.js[c] : m1.vs[c], i2 = vj ? m2.js[c] : m2.vs[c];
                if (i1 == i2) continue;
                const int l1 = (int)in.refseq_len[i1], l2 = (int)in.refseq_len[i2], mn = min(l1, l2);
                const uint8_t *a = in.refseq_arena + in.refseq_off[i1] + (vj ? l1 - mn : 0), *b = in.refseq_arena + in.refseq_off[i2] + (vj ? l2 - mn : 0);
                for (int p = lane; p < mn; p += 32) refdiff += (a[p] != b[p]);
                if (lane == 0) refdiff += abs(l1 - l2);
            }
        }
    }
    pc.refdiff = warp_sum(refdiff);
    pc.diffs = warp_sum(diffs); pc.cd0 = warp_sum(cdc[0]); pc.cd1 = warp_sum(cdc[1]);
    for (int u = 0; u < 2; u++) {
        pc.shares[u] = warp_sum(sh[u]); pc.indeps[u] = warp_sum(in_[u]); pc.indeps_out[u] = warp_sum(io[u]);
        for (int q = 0; q < 4; q++) pc.det[u][q] = warp_sum(det[u][q]);
    }
    pc.lfw = 0; pc.dfw = pc.dc12 = pc.l12 = 0;
    if (reg) { pc.dfw = warp_sum(dfw); pc.dc12 = warp_sum(d12); pc.lfw = c1 - f1; pc.l12 = (f2 - c1) + (f3 - c2); }
}

__device__ __forceinline__ void bytewise_counts(const DevIn &in, const uint4 *t2rows, const uint32_t *c3rows, const DevParams &pr, const EntryMeta &m1, const EntryMeta &m2, PairCounts &pc) {
    bytewise_counts_t<true>(in, t2rows, c3rows, pr, m1, m2, pc);
}
// K6  tier 2, bytewise: one warp per queued pair (already past the CDR3 and meta gates).
__global__ void __launch_bounds__(128) k_tier2_generic(DevIn in, DevTables tb, DevParams pr, DevCounters *ctr, const uint2 *generic_q,
                                                       unsigned long long cap, ejoin_raw_edge_t *acc, uint32_t k_base, uint32_t *parent_min,
                                                       int round2) {
    const unsigned long long q_begin = round2 ? ctr->n_generic_r1 : 0ull;
    unsigned long long nq = ctr->n_generic_q;
    if (nq > cap) nq = cap;
    const uint32_t lane = threadIdx.x & 31;
    for (unsigned long long i = q_begin + (((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < nq;
         i += ((unsigned long long)gridDim.x * blockDim.x) >> 5) {
        const uint2 c = generic_q[i];
        const EntryMeta m1 = tb.meta[c.x], m2 = tb.meta[c.y];      // already ordered k1 < k2
        PairCounts pc;
        bytewise_counts(in, tb.t2, tb.c3, pr, m1, m2, pc);
        if (lane == 0) atomicAdd(&ctr->n_generic, 1ull);
        bool ok = true;
        if (!pr.is_bcr || pr.max_diffs < 1000000) {
            int capd = pr.max_diffs;
            if (!pr.is_bcr && pr.max_diffs_tcr < capd) capd = pr.max_diffs_tcr;
            if (pc.diffs > capd) ok = false;
        }
        int reason = 0, kk = 0, d = 0, err = 0;
        double p1, mult, score;
        if (ok) {
            if (m1.donor_lo != EJOIN_NONE && m2.donor_lo != EJOIN_NONE)
                err = !(m1.donor_lo == m1.donor_hi && m2.donor_lo == m2.donor_hi && m1.donor_lo == m2.donor_lo);
            reason = decide(pr, tb, m1, m2, pc, ctr, &kk, &d, &p1, &mult, &score);
        }
        if (lane == 0) {
            if (reason != 0 && parent_min) atomicMin(&parent_min[c.y], c.x);
            emit_raw(pr, ctr, acc, cap, m1.k + k_base, m2.k + k_base, pc.cd0 + pc.cd1, d, reason, err, reason != 0);
        }
    }
}

// K7  payload: the full PotentialJoin record of each emitted edge.
//   k_payload_fast  one thread per edge, bit-parallel over the t2 rows (same germline, no flags);
//                   the edges it cannot represent are listed for
//   k_payload       one warp per listed edge, bytewise from the ASCII bytes.
__device__ __forceinline__ void write_edge(ejoin_edge_t *out, uint32_t k1, uint32_t k2, const PairCounts &pc, const EntryMeta &m1, int kk, int d,
                                           int err, int reason, double p1, double mult, double score) {
    ejoin_edge_t e;
    memset(&e, 0, sizeof(e));
    e.k1 = k1; e.k2 = k2;
    for (int u = 0; u < pc.nrefs; u++) {
        e.shares[u] = pc.shares[u]; e.indeps[u] = pc.indeps[u];
        for (int q = 0; q < 4; q++) e.shares_details[u][q] = (uint16_t)pc.det[u][q];
    }
    e.nrefs = (uint16_t)pc.nrefs; e.cd = (uint16_t)(pc.cd0 + pc.cd1); e.cd_chain[0] = (uint16_t)pc.cd0; e.cd_chain[1] = (uint16_t)pc.cd1;
    e.diffs = (uint32_t)pc.diffs; e.k = (uint16_t)kk; e.d = (uint16_t)d;
    e.n = 3u * ((uint32_t)m1.len[0] + (uint32_t)m1.len[1]);
    e.err = (uint8_t)err; e.accept_reason = (uint8_t)reason;
    e.p1 = p1; e.mult = mult; e.score = score;
    *out = e;
}

__global__ void __launch_bounds__(128) k_payload_fast(DevIn in, DevTables tb, DevParams pr, DevCounters *ctr, const uint32_t *k_to_s,
                                                      const uint2 *pairs, uint32_t npairs, RangeTab rt, uint32_t n_held, ejoin_edge_t *out, uint32_t *slow_list) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += gridDim.x * blockDim.x) {
        const uint2 kk2 = pairs[i];
        const EntryMeta &m1 = tb.meta[k_to_s[rt_to_local(rt, n_held, kk2.x)]], &m2 = tb.meta[k_to_s[rt_to_local(rt, n_held, kk2.y)]];
        const bool generic = ((m1.flags | m2.flags) & (FLAG_GENERIC0 | FLAG_GENERIC1)) || m1.vs[0] != m2.vs[0] || m1.vs[1] != m2.vs[1] ||
                             m1.js[0] != m2.js[0] || m1.js[1] != m2.js[1];
        if (generic) { slow_list[atomicAdd(&ctr->n_payload_slow, 1u)] = i; continue; }
        PairCounts pc;
        memset(&pc, 0, sizeof(pc));
        pc.nrefs = 1;
        for (int c = 0; c < 2; c++)
            if (m1.vref[c] != m2.vref[c] || m1.dref[c] != m2.dref[c] || m1.jref[c] != m2.jref[c]) pc.nrefs = 2;
        // CDR3 differences per chain
        const uint32_t cl0 = m1.cl[0], total = cl0 + m1.cl[1], W = max(1u, (total + 31) / 32);
        int cdt = 0, cd0 = 0;
        for (uint32_t w = 0; w < W; w++) {
            const uint32_t D = (tb.c3[m1.c3off + w] ^ tb.c3[m2.c3off + w]) | (tb.c3[m1.c3off + W + w] ^ tb.c3[m2.c3off + W + w]);
            cdt += __popc(D);
            cd0 += __popc(D & range_mask((int)w, 0, (int)cl0));
        }
        pc.cd0 = cd0; pc.cd1 = cdt - cd0;
        int f1 = 0, c1 = 0, f2 = 0, c2_ = 0, f3 = 0;
        const bool reg = pr.is_bcr && region_bounds(m1, m2, &f1, &c1, &f2, &c2_, &f3);
        const uint32_t *w1 = (const uint32_t *)(tb.t2 + m1.t2off), *w2 = (const uint32_t *)(tb.t2 + m2.t2off);
        int diffs = 0, dfw = 0, d12 = 0, io = 0, det[4] = { 0, 0, 0, 0 }, blk0 = 0;
        for (int c = 0; c < 2; c++) {
            const int nw = ((int)m1.len[c] + 127) / 128 * 4;
            const int vend = m1.vend[c];                    // same germline and length => same windows for both entries
            const int a1 = m1.cs[c], a2 = m2.cs[c], cl = m1.cl[c];
            for (int w = 0; w < nw; w++) {
                const int o = (blk0 + (w >> 2)) * 12 + (w & 3);
                const uint32_t D = (w1[o] ^ w2[o]) | (w1[o + 4] ^ w2[o + 4]);
                const uint32_t M1 = w1[o + 8], M2 = w2[o + 8];
                const uint32_t S = M1 & M2 & ~D;
                diffs += __popc(D);
                const uint32_t inV = range_mask(w, 0, vend);                 // share in the V window, else in the J window
                det[2 * c] += __popc(S & inV);
                det[2 * c + 1] += __popc(S & ~inV);
                const uint32_t R = range_mask(w, a1, a1 + cl) | range_mask(w, a2, a2 + cl);
                io += __popc(M1 & D & ~R) + __popc(M2 & D & ~R);
                if (reg && c == 0) {
                    dfw += __popc(D & range_mask(w, f1, c1));
                    d12 += __popc(D & (range_mask(w, c1, f2) | range_mask(w, c2_, f3)));
                }
            }
            blk0 += nw / 4;
        }
        const int shares = det[0] + det[1] + det[2] + det[3];
        const int mm = (int)m1.m[0] + (int)m1.m[1] + (int)m2.m[0] + (int)m2.m[1];
        pc.diffs = diffs;
        for (int u = 0; u < 2; u++) {
            pc.shares[u] = shares; pc.indeps[u] = mm - 2 * shares; pc.indeps_out[u] = io;
            for (int q = 0; q < 4; q++) pc.det[u][q] = det[q];
        }
        pc.lfw = 0; pc.dfw = pc.dc12 = pc.l12 = 0;
        if (reg) { pc.dfw = dfw; pc.lfw = c1 - f1; pc.dc12 = d12; pc.l12 = (f2 - c1) + (f3 - c2_); }
        int kk = 0, d = 0, err = 0;
        double p1 = 0, mult = 0, score = 0;
        if (m1.donor_lo != EJOIN_NONE && m2.donor_lo != EJOIN_NONE)
            err = !(m1.donor_lo == m1.donor_hi && m2.donor_lo == m2.donor_hi && m1.donor_lo == m2.donor_lo);
        const int reason = decide(pr, tb, m1, m2, pc, ctr, &kk, &d, &p1, &mult, &score);
        write_edge(&out[i], kk2.x, kk2.y, pc, m1, kk, d, err, reason, p1, mult, score);
    }
}

__global__ void __launch_bounds__(128) k_payload(DevIn in, DevTables tb, DevParams pr, DevCounters *ctr, const uint32_t *k_to_s,
                                                 const uint2 *pairs, const uint32_t *slow_list, RangeTab rt, uint32_t n_held, ejoin_edge_t *out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nslow = ctr->n_payload_slow;
    for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < nslow; j += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t i = slow_list[j];
        const uint2 kk2 = pairs[i];
        const EntryMeta m1 = tb.meta[k_to_s[rt_to_local(rt, n_held, kk2.x)]], m2 = tb.meta[k_to_s[rt_to_local(rt, n_held, kk2.y)]];
        PairCounts pc;
        bytewise_counts(in, tb.t2, tb.c3, pr, m1, m2, pc);
        int kk = 0, d = 0, err = 0;
        double p1 = 0, mult = 0, score = 0;
        if (m1.donor_lo != EJOIN_NONE && m2.donor_lo != EJOIN_NONE)
            err = !(m1.donor_lo == m1.donor_hi && m2.donor_lo == m2.donor_hi && m1.donor_lo == m2.donor_lo);
        int reason = decide(pr, tb, m1, m2, pc, ctr, &kk, &d, &p1, &mult, &score);
        if (lane == 0) write_edge(&out[i], kk2.x, kk2.y, pc, m1, kk, d, err, reason, p1, mult, score);
    }
}

// ---------------------------------------------------------------------------------------------
// K8  collect: phase-A tree edges (one slot per entry, ~0 = none) followed by the phase-B
// accepted edges, as (key = k1<<32|k2, val = flags<<32|d<<16|cd) for the sort.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_collect(const unsigned long long *f0_key, const unsigned long long *f0_val, uint32_t n,
                                                 const ejoin_raw_edge_t *acc, DevCounters *ctr, unsigned long long cap,
                                                 unsigned long long *keys, unsigned long long *vals) {
    // compacted: the sort that follows then handles the live edges only (n_kept of them)
    __shared__ SplitSmem sm;
    unsigned long long nacc = ctr->n_accept;
    if (nacc > cap) nacc = cap;
    const unsigned long long total = (unsigned long long)n + nacc;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x; t0 < total; t0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = t0 + threadIdx.x;
        unsigned long long key = ~0ull, val = 0;
        if (i < n) { key = f0_key[i]; val = f0_val[i]; }
        else if (i < total) {
            const ejoin_raw_edge_t e = acc[i - n];
            if (e.k1 != 0xFFFFFFFFu) {
                key = ((unsigned long long)e.k1 << 32) | e.k2;
                val = ((unsigned long long)e.flags << 32) | ((unsigned long long)e.d << 16) | e.cd;
            }
        }
        unsigned long long o1, o2;
        block_reserve2(sm, key != ~0ull ? 1u : 0u, 0u, &ctr->n_kept, nullptr, o1, o2);
        if (key != ~0ull && o1 < cap) { keys[o1] = key; vals[o1] = val; }
    }
}

__global__ void k_init_rows(uint32_t *parent, unsigned long long *f0_key, uint32_t *cnt, uint8_t *needed, uint8_t *force, uint32_t n,
                            uint8_t needed_init) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = i; f0_key[i] = ~0ull; cnt[i] = 0; needed[i] = needed_init; force[i] = 0; }
}

// Every entry to its phase-A tree root, in one launch: parent[s] <= s always, roots never move while this runs, and a
// concurrent store only replaces a pointer by another ancestor, so every walk ends at the root of its tree.  (Ten
// launches of four hops each used to do this; chains are short — a child hangs off its SMALLEST accepted neighbour.)
__global__ void k_root_all(uint32_t *parent, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = __ldcg(&parent[i]);
    for (;;) { const uint32_t q = __ldcg(&parent[p]); if (q == p) break; p = q; }
    parent[i] = p;
}


// ---------------------------------------------------------------------------------------------
// p1 tables: one block per n (= 3 * L), see dd.cuh.  Phase 1: T_j(k) into scratch;
// phase 2: p1[d][k] = 1 - T_0 - ... - T_{d-1}, clamped at 0 like the reference.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(P1_DCAP) k_p1_tables(const uint32_t *Ls, double *const *tabs, dd_t *scratch) {
    __shared__ dd_t F[P1_KCAP + 1];
    __shared__ dd_t B[2][P1_DCAP];
    const double n = 3.0 * (double)Ls[blockIdx.x];
    double *tab = tabs[blockIdx.x];
    dd_t *T = scratch + (size_t)blockIdx.x * P1_KCAP * P1_DCAP;
    const int j = threadIdx.x;
    if (j == 0) {
        dd_t f = dd_from(1.0);
        F[0] = f;
        for (int u = 0; u < P1_KCAP; u++) { f = p1_f_step(f, u, n); F[u + 1] = f; }
    }
    B[0][j] = dd_from(j == 0 ? 1.0 : 0.0);           // k = 0
    __syncthreads();
    T[j] = j == 0 ? F[0] : dd_from(0.0);
    for (int k = 1; k < P1_KCAP; k++) {
        const dd_t *prev = B[(k - 1) & 1];
        dd_t b = p1_b_step(k, j, prev[j > 0 ? j - 1 : 0], prev[j], n);
        B[k & 1][j] = b;
        T[(size_t)k * P1_DCAP + j] = j <= k ? dd_mul(b, F[k - j]) : dd_from(0.0);
        __syncthreads();
    }
    for (int k = j; k < P1_KCAP; k += P1_DCAP) {
        dd_t p = dd_from(1.0);
        for (int d = 0; d < P1_DCAP; d++) {
            double v = dd_to_double(p);
            tab[(size_t)d * P1_KCAP + k] = v < 0.0 ? 0.0 : v;
            p = dd_sub(p, T[(size_t)k * P1_DCAP + d]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4  sweep: tier 1 — CDR3 XOR/popcount of all pairs of a comparison group against the group's cd cap.
//
// Work item = one tile row (group g, B tile ib): the CTA's 128 threads each own one entry b of the B
// tile (its c3 row in registers) and sweep the A tiles ia = 0..ib in ascending order; A tiles are
// staged in shared memory with 1-D bulk TMA copies, double buffered on two mbarriers.  Inside a group
// sorted position order == info order, so thread b meets its CDR3 survivors a < b in ascending info
// order: the first PHASE_A_K of them go to b's slot array (phase A scores them in that order), the
// rest to the global survivor list (phase B).  Every entry that appears in a survivor pair is flagged
// `needed`: only those get their V..J rows packed (about half of a repertoire never needs them).
// ---------------------------------------------------------------------------------------------
#define PHASE_A_K 6u
// A record of the survivor list: the survivors of entry b among the 32 entries a = abase .. abase + 31 as a bit mask
// (x = abase, y = b, z = mask, w = flags: bit 0 = b gave up in phase A and all of these are scored in round 1).  A clonal
// family meeting itself fills whole masks: 16 bytes then stand for up to 32 pairs (the pair list this replaces was 0.5 GB
// at the 1.6 M-cell input, written by the sweep and read back by the split).  Phase A records single pairs as mask 1.
struct SweepCtx {
    DevCounters *ctr;
    uint4 *cand; unsigned long long cand_cap;
    uint32_t *firstk;      // [n][PHASE_A_K] a positions: segment 0 of every row
    uint32_t *cnt;         // [n] survivors of b in segment 0
    uint32_t *xk;          // [xrow][PHASE_A_K][T1_TILE] a positions: segments q >= 1
    uint32_t *xcnt;        // [xrow][T1_TILE]
    unsigned long long xcap;
    uint8_t *needed;       // [n]
};

#define SWEEP_STAGE 96u        // survivor records staged per warp
__device__ __forceinline__ void sweep_flush(SweepCtx &cx, uint4 *stage, uint32_t &fill) {
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();
    if (fill) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&cx.ctr->n_cand, (unsigned long long)fill);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (uint32_t i = lane; i < fill; i += 32)
            if (base + i < cx.cand_cap) cx.cand[base + i] = stage[i];
    }
    __syncwarp();
    fill = 0;
}

// What a thread does with the survivor mask of its entry b against one block of 32 a's (posA = sorted position of the block's
// first a): needed flags, the first PHASE_A_K survivors of the segment into b's slots, the rest as one 16-byte record on the
// survivor list.  Warp-collective: every lane of the warp calls it with its own mask (0 = nothing).
template <uint32_t STAGE_CAP>
__device__ __forceinline__ void sweep_emit(SweepCtx &cx, uint32_t bits, uint32_t posA, uint32_t posB, uint32_t &mycnt, uint4 *stage, uint32_t &fill,
                                           uint32_t *slot, uint32_t slot_stride) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t col = __reduce_or_sync(0xFFFFFFFFu, bits);      // a's with a survivor partner in this warp
    if (!col) return;
    if ((col >> lane) & 1u) cx.needed[posA + lane] = 1;
    // the first PHASE_A_K survivors of b, in ascending a, go to its slot array
    while (bits && mycnt < PHASE_A_K) {
        const uint32_t aa = __ffs(bits) - 1;
        bits &= bits - 1;
        slot[(size_t)mycnt * slot_stride] = posA + aa;
        mycnt++;
    }
    // the rest: one record per (b, block of 32 a's) on the global survivor list, through the warp's staging buffer (one
    // atomic per STAGE_CAP records and coalesced 16-byte stores)
    const uint32_t have = __ballot_sync(0xFFFFFFFFu, bits != 0);
    if (have) {
        mycnt += __popc(bits);
        const uint32_t totalc = __popc(have);
        if (fill + totalc > STAGE_CAP) sweep_flush(cx, stage, fill);
        if (bits) stage[fill + __popc(have & ((1u << lane) - 1u))] = make_uint4(posA, posB, bits, 0u);
        fill += totalc;
    }
}

template <int W>
__device__ __forceinline__ void row_sweep(SweepCtx &cx, const uint32_t *__restrict__ sA, uint32_t rowsA, bool diag, uint32_t cdmax,
                                          const uint32_t (&bH)[W], const uint32_t (&bL)[W], bool validB, uint32_t posA0, uint32_t posB,
                                          uint32_t &mycnt, uint4 *stage, uint32_t &fill, uint32_t *slot, uint32_t slot_stride) {
    const uint32_t tid = threadIdx.x;
    if (!__any_sync(0xFFFFFFFFu, validB)) return;          // a warp beyond the end of a short B tile
    for (uint32_t a0 = 0; a0 < rowsA; a0 += 32) {
        // on the diagonal tile only a < b counts: a warp whose b's all precede this block is done
        if (diag && a0 > (tid | 31u)) break;
        uint32_t bits = 0;
        const uint32_t na_loop = min(32u, rowsA - a0);
#pragma unroll 4
        for (uint32_t aa = 0; aa < na_loop; aa++) {
            const uint32_t *ra = sA + (a0 + aa) * 2 * W;
            uint32_t cnt = 0;
#pragma unroll
            for (int w = 0; w < W; w++) cnt += __popc((bH[w] ^ ra[w]) | (bL[w] ^ ra[W + w]));
            bits |= (cnt <= cdmax ? 1u : 0u) << aa;
        }
        if (!validB) bits = 0;
        if (diag) {                                  // keep a < b  <=>  a0 + aa < tid
            if (tid <= a0) bits = 0;
            else if (tid - a0 < 32) bits &= (1u << (tid - a0)) - 1u;
        }
        sweep_emit<SWEEP_STAGE>(cx, bits, posA0 + a0, posB, mycnt, stage, fill, slot, slot_stride);
    }
}

template <int W>
__device__ __forceinline__ void row_item(SweepCtx &cx, const GroupInfo &g, uint32_t ib, uint32_t q, const uint32_t *__restrict__ c3,
                                         uint32_t (*sA)[T1_TILE * 2 * T1_WMAX], uint64_t *bar, uint32_t &parity, uint4 *stage, uint32_t &fill) {
    const uint32_t tid = threadIdx.x;
    const uint32_t rowsB = min((uint32_t)T1_TILE, g.n - ib * T1_TILE);
    const bool validB = tid < rowsB;
    const uint32_t posB = g.start + ib * T1_TILE + tid;
    const uint32_t S = g.seg, ia_lo = q * S, ia_hi = min(ia_lo + S - 1, ib);
    // where this segment's first survivors go
    uint32_t *slot = cx.firstk + (size_t)posB * PHASE_A_K, slot_stride = 1;
    unsigned long long xrow = 0;
    bool slots_ok = true;
    if (q) {
        xrow = (unsigned long long)g.xoff + seg_item_off(g.tiles, S, q) + (g.tiles - 1 - ib) - g.tiles;
        slots_ok = xrow < cx.xcap;                       // too small: the host sees n_xrows > cap and retries
        slot = cx.xk + (size_t)(slots_ok ? xrow : 0) * PHASE_A_K * T1_TILE + tid; slot_stride = T1_TILE;
    }
    uint32_t bH[W], bL[W];
    {
        const uint32_t *rb = c3 + g.c3base + (size_t)(ib * T1_TILE + (validB ? tid : 0)) * 2 * W;
#pragma unroll
        for (int w = 0; w < W; w++) { bH[w] = rb[w]; bL[w] = rb[W + w]; }
    }
    uint32_t mycnt = 0;
    auto issue = [&](uint32_t ia) {
        const uint32_t b = (ia - ia_lo) & 1;
        const uint32_t rowsA = min((uint32_t)T1_TILE, g.n - ia * T1_TILE);
        const uint32_t bytes = (rowsA * 2 * W * 4 + 15) & ~15u;
        fence_proxy_async();
        mbar_expect_tx(&bar[b], bytes);
        tma_load_1d(sA[b], c3 + g.c3base + (size_t)ia * T1_TILE * 2 * W, bytes, &bar[b]);
    };
    if (tid == 0) issue(ia_lo);
    for (uint32_t ia = ia_lo; ia <= ia_hi; ia++) {
        const uint32_t b = (ia - ia_lo) & 1;
        if (tid == 0 && ia + 1 <= ia_hi) issue(ia + 1);    // the other buffer was released by the barrier ending the previous step
        mbar_wait(&bar[b], (parity >> b) & 1u);
        parity ^= 1u << b;
        const uint32_t rowsA = min((uint32_t)T1_TILE, g.n - ia * T1_TILE);
        if (slots_ok)
            row_sweep<W>(cx, sA[b], rowsA, ia == ib, g.cdmax, bH, bL, validB, g.start + ia * T1_TILE, posB, mycnt, stage, fill, slot, slot_stride);
        __syncthreads();
    }
    if (q == 0) { if (validB && mycnt) cx.cnt[posB] = mycnt; }
    else if (slots_ok) cx.xcnt[(size_t)xrow * T1_TILE + tid] = validB ? mycnt : 0u;
    if (validB && mycnt) cx.needed[posB] = 1;
}

__global__ void __launch_bounds__(T1_TILE) k_sweep(const uint32_t *__restrict__ c3, const GroupInfo *__restrict__ groups, DevCounters *ctr,
                                                   const uint2 *__restrict__ order, uint4 *cand, unsigned long long cand_cap, uint32_t *firstk,
                                                   uint32_t *cnt, uint32_t *xk, uint32_t *xcnt, unsigned long long xcap, uint8_t *needed) {
    __shared__ __align__(128) uint32_t sA[2][T1_TILE * 2 * T1_WMAX];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t s_item;
    __shared__ uint4 s_stage[T1_TILE / 32][SWEEP_STAGE];
    const uint32_t tid = threadIdx.x;
    uint4 *stage = s_stage[tid >> 5];
    uint32_t fill = 0;                         // pairs in this warp's staging buffer (warp-uniform)
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    __syncthreads();
    uint32_t parity = 0;                       // bit i = phase of bar[i]
    SweepCtx cx;
    cx.ctr = ctr; cx.cand = cand; cx.cand_cap = cand_cap; cx.firstk = firstk; cx.cnt = cnt; cx.xk = xk; cx.xcnt = xcnt; cx.xcap = xcap;
    cx.needed = needed;
    const unsigned long long nitems = ctr->n_items;
    for (;;) {
        if (tid == 0) s_item = (uint32_t)atomicAdd(&ctr->work_next, 1ull);
        __syncthreads();
        const unsigned long long w = s_item;
        __syncthreads();
        if (w >= nitems) break;
        const uint2 it = order[w];             // longest items first (k_item_scatter)
        const GroupInfo g = groups[it.x];
        if (g.mma != MMA_NONE) continue;       // k_sweep_mma's
        const uint32_t q = it.y >> 16, ib = it.y & 0xFFFFu;
        switch (g.W) {
            case 1: row_item<1>(cx, g, ib, q, c3, sA, bar, parity, stage, fill); break;
            case 2: row_item<2>(cx, g, ib, q, c3, sA, bar, parity, stage, fill); break;
            case 3: row_item<3>(cx, g, ib, q, c3, sA, bar, parity, stage, fill); break;
            case 4: row_item<4>(cx, g, ib, q, c3, sA, bar, parity, stage, fill); break;
            case 5: row_item<5>(cx, g, ib, q, c3, sA, bar, parity, stage, fill); break;
            case 6: row_item<6>(cx, g, ib, q, c3, sA, bar, parity, stage, fill); break;
            case 7: row_item<7>(cx, g, ib, q, c3, sA, bar, parity, stage, fill); break;
            default: row_item<8>(cx, g, ib, q, c3, sA, bar, parity, stage, fill); break;
        }
    }
    sweep_flush(cx, stage, fill);
}

// ---------------------------------------------------------------------------------------------
// Tier 1 on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM) for large groups of short CDR3s.
// A base becomes a vertex of the regular tetrahedron in {+1,-1}^3 — v_g . v_h = 3 if g == h, -1 otherwise — so for two rows of one
// group (T bases each)  dot = 3 (T - d) - d = 3T - 4d  and  d <= cdmax  <=>  dot >= thr := 3T - 4 cdmax,  exact in s32; two more K slots
// — (16, 1) in every row, (-thr / 16, -thr % 16) patched into the B tile — put -thr into the contraction, so a survivor is a clear sign bit.  The CTA keeps
// the B tile (128 entries = the M rows of the MMA) in shared memory and sweeps the A tiles of the work item (128 entries = the N
// columns) exactly as k_sweep does; the tiles arrive by 1-D bulk TMA copies in the image the MMA wants (K-major, no swizzle: [K/16][128
// rows][16 bytes]; descriptor LBO = 2048 B between the two 16-byte K chunks of one instruction, SBO = 128 B between groups of 8
// rows), one thread issues the K/32 instructions of a tile pair and commits them to an mbarrier, two accumulator buffers of 128 TMEM
// columns let the MMA of tile i+1 run under the epilogue of tile i, and the epilogue (tcgen05.ld 32x32b.x32) hands thread b = TMEM
// lane the dot products of 32 consecutive a's: the 32-bit survivor mask of (b, block of 32 a's) that sweep_emit takes.
// tools/micro/hamming_mma.cu is the stand-alone prototype of this kernel (same instructions, checked against a popcount kernel).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle: start address, leading (K chunk) and stride (8-row group) byte offsets in 16-byte
// units, descriptor version 1 (Blackwell), layout type 0
__device__ __forceinline__ unsigned long long umma_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, unsigned long long da, unsigned long long db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}

// int8 rows of the tensor-core groups from their c3 rows: one thread per sorted entry, 16-byte stores (32 consecutive entries write 512
// contiguous bytes of one K chunk).  Positions behind the group's base count stay zero.
__global__ void __launch_bounds__(256) k_pack_i8(const uint32_t *__restrict__ c3, const uint32_t *__restrict__ gid_incl, const GroupInfo *__restrict__ groups,
                                                 const DevCounters *ctr, int8_t *__restrict__ tiles) {
    if (ctr->i8_tiles == 0) return;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (uint32_t)ctr->n_valid) return;
    const GroupInfo g = groups[gid_incl[s] - 1];
    if (g.mma == MMA_NONE) return;
    const uint32_t total = g.mma >> 23, slot0 = g.mma & 0x7FFFFFu, W = g.W;
    const uint32_t rel = s - g.start, ib = rel / T1_TILE, row = rel % T1_TILE;
    const uint32_t *rp = c3 + g.c3base + (size_t)rel * 2 * W;
    uint32_t H[3], L[3];
#pragma unroll
    for (int w = 0; w < 3; w++) { H[w] = w < (int)W ? rp[w] : 0u; L[w] = w < (int)W ? rp[W + w] : 0u; }
    int8_t *dst = tiles + (size_t)(slot0 + ib) * MMA_TILE_BYTES + (size_t)row * 16;
    const uint32_t chunks = MMA_KSTEPS(total) * 2;
    for (uint32_t j = 0; j < chunks; j++) {
        uint32_t v[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const uint32_t k = j * 16 + q, p = k / 3, c = k - 3 * p;
            uint32_t x = 0;
            if (p < total) {
                const uint32_t code = (((H[p >> 5] >> (p & 31)) & 1u) << 1) | ((L[p >> 5] >> (p & 31)) & 1u);
                // codes 0..3 -> (1,1,1) (1,-1,-1) (-1,1,-1) (-1,-1,1)
                const bool neg = (code == 1 && c != 0) || (code == 2 && c != 1) || (code == 3 && c != 2);
                x = neg ? 0xFFu : 0x01u;
            }
            if (k == 3 * total) x = 16u;               // the a side of the two threshold slots: with (-thr / 16, -thr % 16) patched into the B tile
            if (k == 3 * total + 1) x = 1u;            // (k_sweep_mma) the accumulator is dot - thr and a survivor is a clear sign bit
            v[q >> 2] |= x << (8 * (q & 3));
        }
        *reinterpret_cast<uint4 *>(dst + (size_t)j * T1_TILE * 16) = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

#define MMA_STAGE 32u          // survivor records staged per warp (shared memory is the tiles')
#define MMA_THREADS 192u       // warps 0-3: epilogue (thread = row b); warp 4 lane 0: MMA issuer; warp 5 lane 0: TMA producer
#define MMA_NA 2u              // shared-memory stages of A tiles
#define MMA_NACC 2u            // accumulator stages in TMEM (128 columns each)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// The three roles walk the same sequence of work items (the cost-sorted list dealt round-robin to the CTAs, the popcount groups' items
// skipped) and of A tiles inside an item, each with its own running tile counter t: stage t % MMA_NA of shared memory, accumulator
// t % MMA_NACC, barrier phase (t / stages) & 1.  Producer: waits until the MMAs that read a stage are complete (bar_sfree), loads the
// next A tile into it.  Issuer: waits for the tile (bar_full) and for the epilogue to have read the accumulator (bar_empty), issues the
// K/32 instructions, commits to bar_sfree and bar_done — it never waits for an MMA to finish, so the next tile pair queues behind the
// running one.  Epilogue: waits bar_done, loads its 128 dot products from TMEM, releases the accumulator (bar_empty, one arrival per
// warp) and only then turns them into survivor masks.
struct MmaItem {
    GroupInfo g; uint32_t q, ib, total, slot0, ksteps, tile_bytes, ia_lo, ia_hi; unsigned long long xrow; bool slots_ok;
};
__device__ __forceinline__ bool mma_item(const GroupInfo *__restrict__ groups, const uint2 *__restrict__ order, unsigned long long w, unsigned long long xcap, MmaItem &m) {
    const uint2 it = order[w];
    m.g = groups[it.x];
    if (m.g.mma == MMA_NONE) return false;                // k_sweep's
    m.q = it.y >> 16; m.ib = it.y & 0xFFFFu;
    m.total = m.g.mma >> 23; m.slot0 = m.g.mma & 0x7FFFFFu;
    m.ksteps = MMA_KSTEPS(m.total); m.tile_bytes = m.ksteps * 32 * T1_TILE;
    const uint32_t S = m.g.seg;
    m.ia_lo = m.q * S; m.ia_hi = min(m.ia_lo + S - 1, m.ib);
    m.xrow = 0; m.slots_ok = true;
    if (m.q) {
        m.xrow = (unsigned long long)m.g.xoff + seg_item_off(m.g.tiles, S, m.q) + (m.g.tiles - 1 - m.ib) - m.g.tiles;
        m.slots_ok = m.xrow < xcap;                       // too small: the host sees n_xrows > cap and retries (every role skips the item)
    }
    return m.slots_ok;
}
__global__ void __launch_bounds__(MMA_THREADS, 2) k_sweep_mma(const int8_t *__restrict__ tiles, const GroupInfo *__restrict__ groups, DevCounters *ctr,
                                                           const uint2 *__restrict__ order, uint4 *cand, unsigned long long cand_cap, uint32_t *firstk,
                                                           uint32_t *cnt, uint32_t *xk, uint32_t *xcnt, unsigned long long xcap, uint8_t *needed) {
    if (ctr->i8_tiles == 0) return;                       // no tensor-core group in this input
    extern __shared__ __align__(128) uint8_t mma_smem[];  // B tile | A stages
    __shared__ __align__(8) uint64_t bar_full[MMA_NA], bar_sfree[MMA_NA], bar_done[MMA_NACC], bar_empty[MMA_NACC], bar_bfull, bar_bfree, bar_bpatch;
    __shared__ uint32_t s_tmem;
    __shared__ uint4 s_stage[T1_TILE / 32][MMA_STAGE];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *sB = mma_smem;
    if (tid == 0) {
        for (uint32_t i = 0; i < MMA_NA; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_sfree[i], 1); }
        for (uint32_t i = 0; i < MMA_NACC; i++) { mbar_init(&bar_done[i], 1); mbar_init(&bar_empty[i], T1_TILE / 32); }
        mbar_init(&bar_bfull, 1); mbar_init(&bar_bfree, 1); mbar_init(&bar_bpatch, T1_TILE / 32);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(MMA_NACC * 128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // instruction descriptor: D = s32 (bits 4-5 = 2), A and B signed 8 bit (bits 7-9, 10-12 = 1), both K-major, N = 128 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned long long nitems = ctr->n_items;
    MmaItem m;
    if (warp == 5) {                                      // ---- TMA producer ----
        if (lane == 0) {
            uint32_t t = 0, item_no = 0;
            for (unsigned long long w = blockIdx.x; w < nitems; w += gridDim.x) {
                if (!mma_item(groups, order, w, xcap, m)) continue;
                const int8_t *gt = tiles + (size_t)m.slot0 * MMA_TILE_BYTES;
                mbar_wait(&bar_bfree, (item_no & 1u) ^ 1u);             // the MMAs of the previous item are done with the B tile
                fence_proxy_async();
                mbar_expect_tx(&bar_bfull, m.tile_bytes);
                tma_load_1d(sB, gt + (size_t)m.ib * MMA_TILE_BYTES, m.tile_bytes, &bar_bfull);
                for (uint32_t ia = m.ia_lo; ia <= m.ia_hi; ia++, t++) {
                    const uint32_t st = t % MMA_NA;
                    mbar_wait(&bar_sfree[st], ((t / MMA_NA) & 1u) ^ 1u);
                    mbar_expect_tx(&bar_full[st], m.tile_bytes);
                    tma_load_1d(mma_smem + (size_t)(1 + st) * MMA_TILE_BYTES, gt + (size_t)ia * MMA_TILE_BYTES, m.tile_bytes, &bar_full[st]);
                }
                item_no++;
            }
        }
        __syncwarp();
    } else if (warp == 4) {                               // ---- MMA issuer ----
        if (lane == 0) {
            uint32_t t = 0, item_no = 0;
            for (unsigned long long w = blockIdx.x; w < nitems; w += gridDim.x) {
                if (!mma_item(groups, order, w, xcap, m)) continue;
                mbar_wait(&bar_bpatch, item_no & 1u);                    // the B tile has landed and carries its threshold slots
                for (uint32_t ia = m.ia_lo; ia <= m.ia_hi; ia++, t++) {
                    const uint32_t st = t % MMA_NA, ac = t % MMA_NACC;
                    mbar_wait(&bar_full[st], (t / MMA_NA) & 1u);
                    mbar_wait(&bar_empty[ac], ((t / MMA_NACC) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sB), b_addr = smem_u32(mma_smem + (size_t)(1 + st) * MMA_TILE_BYTES);
                    for (uint32_t k = 0; k < m.ksteps; k++)
                        umma_i8(tmem + ac * 128, umma_smem_desc(a_addr + k * 2 * T1_TILE * 16, T1_TILE * 16, 128),
                                umma_smem_desc(b_addr + k * 2 * T1_TILE * 16, T1_TILE * 16, 128), idesc, k ? 1u : 0u);
                    umma_commit(&bar_sfree[st]);
                    umma_commit(&bar_done[ac]);
                }
                umma_commit(&bar_bfree);
                item_no++;
            }
        }
        __syncwarp();
    } else {                                              // ---- epilogue: thread = entry b of the B tile = TMEM lane ----
        uint4 *stage = s_stage[warp];
        uint32_t fill = 0, t = 0, item_no = 0;
        SweepCtx cx;
        cx.ctr = ctr; cx.cand = cand; cx.cand_cap = cand_cap; cx.firstk = firstk; cx.cnt = cnt; cx.xk = xk; cx.xcnt = xcnt; cx.xcap = xcap;
        cx.needed = needed;
        for (unsigned long long w = blockIdx.x; w < nitems; w += gridDim.x) {
            if (!mma_item(groups, order, w, xcap, m)) continue;
            const GroupInfo &g = m.g;
            const int thr = (int)(3 * m.total) - 4 * (int)g.cdmax;
            const uint32_t rowsB = min((uint32_t)T1_TILE, g.n - m.ib * T1_TILE);
            const bool validB = tid < rowsB;
            const uint32_t posB = g.start + m.ib * T1_TILE + tid;
            uint32_t *slot = cx.firstk + (size_t)posB * PHASE_A_K, slot_stride = 1;
            if (m.q) { slot = cx.xk + (size_t)m.xrow * PHASE_A_K * T1_TILE + tid; slot_stride = T1_TILE; }
            uint32_t mycnt = 0;
            {   // the threshold rides in the contraction: the b side of the two threshold slots, into this thread's row of the B tile
                mbar_wait(&bar_bfull, item_no & 1u);
                const uint32_t k0 = 3 * m.total, k1 = 3 * m.total + 1;
                sB[(k0 / 16) * T1_TILE * 16 + tid * 16 + (k0 % 16)] = (uint8_t)(int8_t)(-(thr / 16));
                sB[(k1 / 16) * T1_TILE * 16 + tid * 16 + (k1 % 16)] = (uint8_t)(int8_t)(-(thr % 16));
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_bpatch);
                item_no++;
            }
            for (uint32_t ia = m.ia_lo; ia <= m.ia_hi; ia++, t++) {
                const uint32_t ac = t % MMA_NACC;
                mbar_wait(&bar_done[ac], (t / MMA_NACC) & 1u);
                tc_fence_after();
                const bool diag = ia == m.ib;
                const uint32_t rowsA = min((uint32_t)T1_TILE, g.n - ia * T1_TILE);
                uint32_t nblk = (rowsA + 31) / 32;
                if (diag) nblk = min(nblk, ((tid | 31u) >> 5) + 1u);     // warp-uniform: blocks that can hold an a < b
                if (!__any_sync(0xFFFFFFFFu, validB)) nblk = 0;          // a warp beyond the end of a short B tile
                uint32_t r[4][32];
                const uint32_t trow = tmem + ((warp * 32u) << 16) + ac * 128;
                if (nblk > 0) tmem_ld32_nowait(trow, r[0]);
                if (nblk > 1) tmem_ld32_nowait(trow + 32, r[1]);
                if (nblk > 2) tmem_ld32_nowait(trow + 64, r[2]);
                if (nblk > 3) tmem_ld32_nowait(trow + 96, r[3]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[ac]);              // this warp has its 32 rows of the accumulator in registers
#pragma unroll
                for (uint32_t blk = 0; blk < 4; blk++) {
                    if (blk >= nblk) break;
                    const uint32_t a0 = blk * 32;
                    // the accumulator is dot - thr: bit j = its sign bit, gathered with one funnel shift per element (four independent chains of
                    // eight); a survivor is a CLEAR sign bit
                    uint32_t part[4] = { 0, 0, 0, 0 };
#pragma unroll
                    for (int j = 7; j >= 0; j--)
#pragma unroll
                        for (int c = 0; c < 4; c++) part[c] = __funnelshift_l(r[blk][8 * c + j], part[c], 1);
                    uint32_t bits = ~(part[0] | (part[1] << 8) | (part[2] << 16) | (part[3] << 24));
                    if (rowsA - a0 < 32) bits &= (1u << (rowsA - a0)) - 1u;
                    if (!validB) bits = 0;
                    if (diag) {                                           // keep a < b
                        if (tid <= a0) bits = 0;
                        else if (tid - a0 < 32) bits &= (1u << (tid - a0)) - 1u;
                    }
                    sweep_emit<MMA_STAGE>(cx, bits, g.start + ia * T1_TILE + a0, posB, mycnt, stage, fill, slot, slot_stride);
                }
            }
            if (m.q == 0) { if (validB && mycnt) cx.cnt[posB] = mycnt; }
            else cx.xcnt[(size_t)m.xrow * T1_TILE + tid] = validB ? mycnt : 0u;
            if (validB && mycnt) cx.needed[posB] = 1;
        }
        sweep_flush(cx, stage, fill);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(MMA_NACC * 128u) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Phase A (lazy exact scoring): one thread per entry b.  b fully scores its first
// survivors in ascending a until one is accepted: that a = m(b) is b's smallest accepted smaller
// neighbour, and (m(b), b) is an edge of the lexicographically-first spanning forest that join_core's
// skip-if-same-class scan would keep.  Survivors after m(b) are only recorded for phase B.  A b whose
// PHASE_A_K slots are all rejected is marked `force`:
// all its remaining survivors are scored in phase B round 1.
// ---------------------------------------------------------------------------------------------
// The slots of entry s (sorted position): segment 0 in firstk/cnt, segments q >= 1 in the extra rows.
struct SlotView {
    const uint32_t *firstk, *cnt, *xk, *xcnt;
    const uint32_t *gid_incl;
    const GroupInfo *groups;
    uint32_t stride, phase;
};
struct RowSlots {
    uint32_t s, t, nseg;
    unsigned long long xbase;      // extra row of segment q = xbase + seg_item_off(T, S, q)
    uint32_t T, S;
    bool owned;
    __device__ __forceinline__ void init(const SlotView &sv, uint32_t s_) {
        s = s_;
        const GroupInfo g = sv.groups[sv.gid_incl[s] - 1];
        const uint32_t rel = s - g.start, ib = rel / T1_TILE;
        t = rel % T1_TILE; T = g.tiles; S = g.seg;
        nseg = ib / S + 1;
        xbase = (unsigned long long)g.xoff + (T - 1 - ib) - T;
        owned = g.n >= 2 && g.cdmax != 0xFFFF &&                                              // a group without work has no slots at all
                !(sv.stride > 1 && (g.work_off + (T - 1 - ib)) % sv.stride != sv.phase);     // same rule as item_decode
    }
    __device__ __forceinline__ uint32_t count(const SlotView &sv, uint32_t q) const {
        return q == 0 ? sv.cnt[s] : sv.xcnt[(size_t)(xbase + seg_item_off(T, S, q)) * T1_TILE + t];
    }
    __device__ __forceinline__ uint32_t slot(const SlotView &sv, uint32_t q, uint32_t i) const {
        return q == 0 ? sv.firstk[(size_t)s * PHASE_A_K + i] : sv.xk[((size_t)(xbase + seg_item_off(T, S, q)) * PHASE_A_K + i) * T1_TILE + t];
    }
};

// Outcome of one entry's walk over its slots (ascending a: segment by segment, slot by slot).
struct WalkOut {
    uint32_t emit0;        // segment-0 slots that go to the survivor list (bit i)
    uint32_t xe_start;     // every slot of the segments q >= 1 with linear index q*K+i >= xe_start goes to the list
    uint32_t n_emit_x;     // how many those are
    bool forced;           // the listed slots carry the forced flag
};

__global__ void __launch_bounds__(128, 8) k_phase_a(DevIn in, DevTables tb, DevParams pr, DevCounters *ctr, SlotView sv, uint8_t *force,
                                                 uint32_t *parent, unsigned long long *f0_key, unsigned long long *f0_val, uint4 *cand,
                                                 unsigned long long cand_cap) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long evals = 0, bytes = 0;
    WalkOut wo;
    wo.emit0 = 0; wo.xe_start = 0xFFFFFFFFu; wo.n_emit_x = 0; wo.forced = false;
    uint32_t defer_li = 0xFFFFFFFFu;           // linear slot index of a pair left to round 1 (bytewise path), see below
    RowSlots rs;
    rs.nseg = 0; rs.owned = false;
    if (s < (uint32_t)ctr->n_valid) {
        rs.init(sv, s);
        if (rs.owned) {
            bool found = false, forced = false, deferred = false;
            uint32_t rejects = 0;
            for (uint32_t q = 0; q < rs.nseg; q++) {
                const uint32_t cq = rs.count(sv, q), nq = min(cq, PHASE_A_K);
                uint32_t i = 0;
                if (!found && !forced && !deferred) {
                    for (; i < nq; i++) {
                        const uint32_t a = rs.slot(sv, q, i);
                        int cd, d, err;
                        bool scored;
                        const int r = evaluate_thread(in, tb, pr, ctr, a, s, &cd, &d, &err, &scored, &bytes);
                        evals += scored;
                        if (r < 0) {
                            // A pair for the bytewise path (has_del / non-ACGT / different germlines; 3 % of the pairs).  The walk
                            // stops here: this slot and the ones behind it in the same segment go to round 1 of phase B with the
                            // forced flag — every smaller survivor of b has been scored and rejected, so the smallest of them that
                            // round 1 accepts IS b's smallest accepted neighbour, and atomicMin makes it the tree edge
                            // (k_forced_resolve) — and the slots of the later segments are recorded like any others.  If none is
                            // accepted b stays a singleton tree and round 2 scores what it recorded.
                            // (Round 1 used to park such walks for a warp-per-entry continuation kernel: 0.35 ms of latency chain.)
                            deferred = true;
                            defer_li = q * PHASE_A_K + i;
                            break;
                        }
                        if (r > 0) {
                            found = true;
                            parent[s] = a;
                            f0_key[s] = ((unsigned long long)tb.meta[a].k << 32) | tb.meta[s].k;
                            f0_val[s] = ((unsigned long long)((uint32_t)r | ((uint32_t)err << 8)) << 32) | ((unsigned long long)(uint32_t)min(d, 65535) << 16) | (uint32_t)cd;
                            i++;
                            break;
                        }
                        rejects++;
                        if (rejects >= PHASE_A_K && (i + 1 < nq || cq > PHASE_A_K || q + 1 < rs.nseg)) { forced = true; i++; break; }
                    }
                    if (!found && !forced && !deferred && cq > PHASE_A_K) forced = true;     // this segment's listed survivors come before the next segment's slots
                }
                // slots i..nq-1 of this segment were not scored: they are recorded for phase B
                if (q == 0) {
                    for (; i < nq; i++) {
                        if (found && parent[rs.slot(sv, 0, i)] == parent[s]) continue;     // hangs off the same parent: same tree, phase B would skip it
                        wo.emit0 |= 1u << i;
                    }
                } else if (i < nq) {
                    wo.xe_start = min(wo.xe_start, q * PHASE_A_K + i);
                    wo.n_emit_x += nq - i;
                }
            }
            if (forced) force[s] = 1;
            wo.forced = forced;
        }
    }
    {
        // one reservation per warp for the recorded survivors
        const uint32_t lane = threadIdx.x & 31;
        const uint32_t cntl = __popc(wo.emit0) + wo.n_emit_x;
        uint32_t incl = cntl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += v; }
        const uint32_t totalc = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (totalc) {
            unsigned long long w = 0;
            if (lane == 0) w = atomicAdd(&ctr->n_cand, (unsigned long long)totalc);
            w = __shfl_sync(0xFFFFFFFFu, w, 0) + (incl - cntl);
            const uint32_t fl = wo.forced ? 1u : 0u;
            uint32_t e = wo.emit0;
            while (e) {
                const uint32_t i = __ffs(e) - 1;
                e &= e - 1;
                if (w < cand_cap) cand[w] = make_uint4(rs.slot(sv, 0, i), s, 1u, (i >= defer_li && defer_li < PHASE_A_K) ? 1u : fl);
                w++;
            }
            if (wo.n_emit_x) {
                for (uint32_t q = max(1u, wo.xe_start / PHASE_A_K); q < rs.nseg; q++) {
                    const uint32_t nq = min(rs.count(sv, q), PHASE_A_K);
                    for (uint32_t i = 0; i < nq; i++) {
                        if (q * PHASE_A_K + i < wo.xe_start) continue;
                        if (w < cand_cap) cand[w] = make_uint4(rs.slot(sv, q, i), s, 1u, (q * PHASE_A_K + i >= defer_li && q == defer_li / PHASE_A_K) ? 1u : fl);
                        w++;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { evals += __shfl_xor_sync(0xFFFFFFFFu, evals, o); bytes += __shfl_xor_sync(0xFFFFFFFFu, bytes, o); }
    if ((threadIdx.x & 31) == 0 && evals) { atomicAdd(&ctr->pairs_tier2, evals); atomicAdd(&ctr->t2_alg_bytes, bytes); }
}

// ---------------------------------------------------------------------------------------------
// EqClass: GPU union-find for finish_join (App. A.7).  Roots are always hooked under the smaller
// index with a CAS, so the root of a class is its smallest member — the label the host side turns
// back into an EquivRel.  Loads bypass L1 (the forest is being rewritten by other CTAs).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t uf_find(uint32_t *p, uint32_t x) {
    for (;;) {
        uint32_t px = __ldcg(&p[x]);
        if (px == x) return x;
        uint32_t gp = __ldcg(&p[px]);
        if (gp != px) p[x] = gp;         // path halving; a stale value is still an ancestor
        x = px;
    }
}
__device__ __forceinline__ void uf_union(uint32_t *p, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find(p, a); b = uf_find(p, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }
        if (atomicCAS(&p[a], a, b) == a) return;
    }
}
__global__ void k_uf_init(uint32_t *label, uint32_t n, uint32_t *first, uint32_t n_exact) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) label[i] = i;
    if (i < n_exact) first[i] = 0xFFFFFFFFu;
}
__global__ void k_uf_first(const uint32_t *exact_id, uint32_t n, uint32_t n_exact, uint32_t *first) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { uint32_t x = exact_id[k]; if (x < n_exact) atomicMin(&first[x], k); }
}
// join the emitted edges, then the info entries that share an exact subclonotype (PREHISTORY:14747)
__global__ void k_uf_union(uint32_t *label, const uint2 *pairs, uint32_t npairs, uint32_t k_base, const uint32_t *exact_id, uint32_t n,
                           uint32_t n_exact, const uint32_t *first) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npairs) uf_union(label, pairs[i].x - k_base, pairs[i].y - k_base);
    if (i < n) {
        uint32_t x = exact_id[i];
        if (x < n_exact) { uint32_t f = first[x]; if (f != i) uf_union(label, f, i); }
    }
}
// Final labels go to a SEPARATE array: other threads are still path-halving through `label`, and a
// halving store (an ancestor, not necessarily the root) must not overwrite a final label.
__global__ void __launch_bounds__(256) k_uf_flatten(uint32_t *label, uint32_t n, DevCounters *ctr, uint32_t *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long isroot = 0;
    if (i < n) { uint32_t r = uf_find(label, i); isroot = r == i; out[i] = r; }
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    unsigned long long s = BR(tmp).Sum(isroot);
    if (threadIdx.x == 0 && s) atomicAdd(&ctr->n_classes, s);
}

// ---------------------------------------------------------------------------------------------
// Finishing the forest on the device (single GPU / contiguous shards).  The phase-A trees are
// certified forest edges; the accepted cross-tree edges from phase B are reduced to the forest
// edges among them by Boruvka rounds over the contracted trees: every component takes the
// lightest (smallest (k1,k2)) edge leaving it — such an edge is in the unique minimum spanning
// forest (cut property) — and the components are merged; repeat until no edge leaves a component.
// `parent` (after the pointer doubling every entry points at its tree root) is the union-find.
// ---------------------------------------------------------------------------------------------
#define ACC_DEAD 0xFFFFFFFFu
#define ACC_SELECTED 0x40000000u     // in ejoin_raw_edge_t.flags

__global__ void k_fill_u64(unsigned long long *p, unsigned long long v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// All rounds in ONE cooperative launch, two grid-wide barriers per round: (1) every live edge finds its two components and bids
// for them (64-bit atomicMin of its (k1,k2) key); (2) the winning edges are selected and their components merged.  The bids
// alternate between two tables (`best2` holds both, nent entries each): while step (2) of round r reads table r & 1, every live
// edge clears its two components' entries in the other table, which is the one round r + 1 bids into.  A component that is bid
// for in round r + 1 is a round-r root that either stayed alone (its own live edge cleared it) or is an end of a selected edge
// (which cleared it), so its entry is clean; entries of absorbed components are never read again.  This took the third barrier
// (a clearing pass) out of every round.  Stops when a round finds no edge between two components.
__global__ void __launch_bounds__(256) k_boruvka(ejoin_raw_edge_t *acc, DevCounters *ctr, unsigned long long cap, const uint32_t *k_to_s,
                                                 uint32_t *parent, unsigned long long *best2, uint2 *ends, uint32_t nent) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    unsigned long long n = ctr->n_accept;
    if (n > cap) n = cap;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x, first = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned int round = 1; round < 4096; round++) {
        unsigned long long *best = best2 + (size_t)(round & 1u) * nent, *other = best2 + (size_t)((round + 1u) & 1u) * nent;
        for (unsigned long long i = first; i < n; i += stride) {
            const ejoin_raw_edge_t e = acc[i];
            if (e.k1 == ACC_DEAD || (e.flags & ACC_SELECTED)) continue;
            const uint32_t ca = uf_find(parent, k_to_s[e.k1]), cb = uf_find(parent, k_to_s[e.k2]);
            if (ca == cb) { acc[i].k1 = ACC_DEAD; continue; }       // both ends in one component: heaviest edge of a cycle
            const unsigned long long key = ((unsigned long long)e.k1 << 32) | e.k2;
            atomicMin(&best[ca], key);
            atomicMin(&best[cb], key);
            ends[i] = make_uint2(ca, cb);
            *(volatile unsigned int *)&ctr->bor_active = round;
        }
        grid.sync();
        if (*(volatile unsigned int *)&ctr->bor_active != round) break;      // nothing left to merge (same answer in every thread)
        for (unsigned long long i = first; i < n; i += stride) {
            const ejoin_raw_edge_t e = acc[i];
            if (e.k1 == ACC_DEAD || (e.flags & ACC_SELECTED)) continue;
            const uint2 c = ends[i];
            const unsigned long long key = ((unsigned long long)e.k1 << 32) | e.k2;
            other[c.x] = ~0ull; other[c.y] = ~0ull;                  // next round's table, clean where this round's components are
            if (best[c.x] == key || best[c.y] == key) {
                acc[i].flags = e.flags | ACC_SELECTED;
                uf_union(parent, c.x, c.y);
            }
        }
        grid.sync();
    }
    if (first == 0) ctr->bor_active = 0;
}
__global__ void k_comp_size(uint32_t *parent, const DevCounters *ctr, uint32_t *size) {       // n_valid read on the device: no host round trip before this launch
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < (uint32_t)ctr->n_valid) atomicAdd(&size[uf_find(parent, s)], 1u);
}
// Final edge list: tree edges + selected cross edges, minus the two-cell rule (help1.rs:349-353):
// an orbit of exactly two entries whose exact subclonotypes total two cells keeps its join only if
// EASY or cd <= d.
__global__ void __launch_bounds__(256) k_final_collect(DevIn in, const unsigned long long *f0_key, const unsigned long long *f0_val, uint32_t n,
                                                       const ejoin_raw_edge_t *acc, DevCounters *ctr, unsigned long long cap,
                                                       const uint32_t *k_to_s, uint32_t k_base, uint32_t *parent, const uint32_t *size, int easy,
                                                       unsigned long long *keys, unsigned long long *vals) {
    // compacted: the sort that follows then handles the emitted joins only (n_kept of them)
    __shared__ SplitSmem sm;
    unsigned long long nacc = ctr->n_accept;
    if (nacc > cap) nacc = cap;
    const unsigned long long total = (unsigned long long)n + nacc;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x; t0 < total; t0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = t0 + threadIdx.x;
        unsigned long long key = ~0ull, val = 0;
        if (i < n) { key = f0_key[i]; val = f0_val[i]; }
        else if (i < total) {
            const ejoin_raw_edge_t e = acc[i - n];
            if (e.k1 != ACC_DEAD && (e.flags & ACC_SELECTED)) {
                key = ((unsigned long long)e.k1 << 32) | e.k2;
                val = ((unsigned long long)(e.flags & ~ACC_SELECTED) << 32) | ((unsigned long long)e.d << 16) | e.cd;
            }
        }
        if (key != ~0ull) {
            const uint32_t k1 = (uint32_t)(key >> 32) - k_base, k2 = (uint32_t)key - k_base;
            if (size[uf_find(parent, k_to_s[k1])] == 2) {
                const uint32_t cells = in.ncells[in.exact_id[k1]] + in.ncells[in.exact_id[k2]];
                const uint32_t cd = (uint32_t)(val & 0xFFFF), d = (uint32_t)((val >> 16) & 0xFFFF);
                if (cells == 2 && !easy && !(cd <= d)) key = ~0ull;
            }
        }
        unsigned long long o1, o2;
        block_reserve2(sm, key != ~0ull ? 1u : 0u, 0u, &ctr->n_kept, nullptr, o1, o2);
        if (key != ~0ull && o1 < cap) { keys[o1] = key; vals[o1] = val; }
    }
}
__global__ void k_keys_to_pairs(const unsigned long long *keys, uint32_t ne, RangeTab rt, uint2 *pairs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ne) pairs[i] = make_uint2(rt_to_global(rt, (uint32_t)(keys[i] >> 32)), rt_to_global(rt, (uint32_t)keys[i]));   // local -> info indices
}
__global__ void k_edge_pairs(const ejoin_edge_t *edges, uint32_t ne, uint2 *pairs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ne) pairs[i] = make_uint2(edges[i].k1, edges[i].k2);
}
// rank 0 after the gather: the ranks' lists are each in raw_joins order, their union is put in order by one key sort
__global__ void k_edge_keys(const ejoin_edge_t *edges, uint32_t ne, unsigned long long *keys, uint32_t *idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ne) { keys[i] = ((unsigned long long)edges[i].k1 << 32) | edges[i].k2; idx[i] = i; }
}
__global__ void k_gather_edges(const ejoin_edge_t *src, const uint32_t *idx, uint32_t ne, ejoin_edge_t *dst) {
    static_assert(sizeof(ejoin_edge_t) % 8 == 0, "ejoin_edge_t is moved as 8-byte words");
    const uint32_t W = sizeof(ejoin_edge_t) / 8;
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long i = t / W, w = t % W;
    if (i < ne) ((unsigned long long *)(dst + i))[w] = ((const unsigned long long *)(src + idx[i]))[w];
}
// Stride mode, rank 0 after the gather: the ranks' accepted edges as (key, val) pairs -> the edge array the forest kernels work on,
// every entry its own component again.
__global__ void k_kv_to_acc(const unsigned long long *keys, const unsigned long long *vals, unsigned long long ne, ejoin_raw_edge_t *acc) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    ejoin_raw_edge_t e;
    e.k1 = (uint32_t)(keys[i] >> 32); e.k2 = (uint32_t)keys[i];
    e.cd = (uint16_t)(vals[i] & 0xFFFF); e.d = (uint16_t)((vals[i] >> 16) & 0xFFFF);
    e.flags = (uint32_t)(vals[i] >> 32) & ~ACC_SELECTED;
    acc[i] = e;
}
__global__ void k_iota_parent(uint32_t *parent, unsigned long long *f0_key, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = i; f0_key[i] = ~0ull; }
}
__global__ void k_set_accept(DevCounters *ctr, unsigned long long n) { ctr->n_accept = n; ctr->n_kept = 0; ctr->bor_active = 0; }
