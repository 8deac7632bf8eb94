// ejoin_dev.cuh — device-side data layout and kernels of the B200 clonotype join.
//
// Path restated: join_exacts -> join_core -> join_one (enclone_ranger@cbd15ef, absent from
// /root/reference; spec: enclone_help/src/help1.rs:270-363).
// Layout in HBM:
//   c3 rows   per comparison group (same V..J lengths, same CDR3 lengths): both CDR3s
//             concatenated, 1 bit-plane pair per 32 bases: [H0..H(W-1) | L0..L(W-1)] u32 words.
//             base code = 2*H + L, A,C,G,T = 0,1,2,3.  diff mask of 32 bases = (H^H')|(L^L').
//   t2 rows   per entry, per chain, per 128-base block: three 128-bit rows
//             [H x4 | L x4 | M x4]; M = 1 where the base differs from the entry's own germline
//             inside the V window [0,|v|-15) or the J window [n-(|j|-15), n)
//             (windows: enclone_tail/src/print_clonotypes/print_utils5.rs:78-81).
//   meta      128-byte record per entry (ids, lengths, region bounds, mutation counts, flags).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/enclone_join.h"
#include "dd.cuh"

#define T1_TILE 128          // entries per tier-1 tile side; CTA = 128 threads
#define T1_WMAX 8            // 32-base words per plane of the concatenated CDR3s (<= 256 nt)
#define C3_MAX_TOTAL (32 * T1_WMAX)
#define FLAG_GENERIC0 1u     // chain 0 needs the bytewise path (has_del, non-ACGT, overlapping windows)
#define FLAG_GENERIC1 2u
#define ERR_CDR3_TOO_LONG 1u
#define ERR_CDR3_BAD_CHAR 2u
#define ERR_BAD_REFSEQ 4u
#define ERR_K_TOO_LARGE 8u
#define ERR_LEN_TOO_LARGE 16u
#define ERR_ARENA_RANGE 128u      // an entry's bytes / words lie outside the uploaded slice of their arena
#define ERR_TABLES_MISSING 256u    // a scored pair needed a p1 / multiplier table the handle has not built yet (optimistic run)
#define ERR_BAD_EXACT 64u         // exact_id[k] >= n_exact
#define ERR_GROUP_TOO_LARGE 32u   // a comparison group of more than 65535 tiles (8.3 M entries with equal lengths)

struct __align__(16) EntryMeta {
    uint32_t k;              // info index
    uint32_t t2off;          // t2 row offset, 16-byte units
    uint32_t c3off;          // c3 row offset, u32 words
    uint32_t donor_lo, donor_hi;
    uint32_t vref[2], jref[2], dref[2];
    uint32_t vs[2], js[2];
    uint32_t vname[2];
    uint32_t c_light;
    uint16_t len[2], cs[2], cl[2];
    uint16_t hcomp, fr1, cdr1, fr2, cdr2, fr3;
    uint16_t m[2];           // germline differences inside the windows, per chain
    uint16_t vend[2], jstart[2];
    uint16_t flags;
    uint16_t pad[9];
};
static_assert(sizeof(EntryMeta) == 128, "EntryMeta must be 128 bytes");

// Tier-1 work decomposition.  A group of T tiles (of T1_TILE entries) is a lower-triangular T x T array
// of tile pairs.  Row ib (its B tile) meets the A tiles 0..ib; the row is cut into segments of `seg`
// A tiles, so that no work item is longer than `seg` tile steps (a row of a 4,000-entry group would
// otherwise run 30+ steps on one CTA while the rest of the GPU idles).  Work items are enumerated
// segment-major: segment q holds the rows ib = T-1 .. q*seg, i.e. T - q*seg items;
//   item_index(q, ib) = q*T - seg*q*(q-1)/2 + (T-1-ib),       items(T) = Q*T - seg*Q*(Q-1)/2,  Q = ceil(T/seg).
// Segment 0 of a row keeps its first survivors in the per-entry slot array; segments q >= 1 use an
// "extra row" of slots, numbered xoff + item_index - T.
#define SEG_MIN 8u           // A tiles per segment (small groups) on a big input; large groups: ceil(T / SEG_QMAX).  Small inputs (a
                             // rank's share of a sharded join) use 4 or 2, so that there are still more work items than CTA slots
// Tier 1 on the tensor cores (k_sweep_mma; tools/micro/hamming_mma.cu is the stand-alone prototype): a group whose two CDR3s total
// at most MMA_MAX_BASES bases and that has at least `mma_min_tiles` tiles gets its rows a second time as int8 tetrahedron vectors
// (3 values per base, then the a side of two threshold slots: 16, 1), one shared-memory image of MMA_TILE_BYTES per tile of 128 entries.
#define MMA_NONE 0xFFFFFFFFu
#define MMA_MAX_BASES 95u                           // 3 * 95 + 2 threshold slots <= MMA_KMAX
#define MMA_KMAX 288u                               // 3 values per base + 2 threshold slots, rounded up to a multiple of 32
#define MMA_KSTEPS(total) ((3u * (total) + 2u + 31u) / 32u)   // tcgen05.mma instructions (K = 32 each) per tile pair
#define MMA_TILE_BYTES (T1_TILE * MMA_KMAX)         // [K/16 chunks][128 rows][16 bytes]; a group with fewer bases uses a prefix
#define MMA_MIN_TILES_DEFAULT 128u                     // 16,384 entries: below that the int8 pack and the second kernel cost more than the popcount sweep of the group (C3: +0.15 ms with 16)
#define SEG_QMAX 32u         // at most this many segments per row
struct GroupInfo {           // one comparison group = one run of equal sort keys
    uint32_t start;          // first sorted position
    uint32_t n;              // entries
    uint32_t c3base;         // u32 word offset of the group's c3 rows
    uint32_t work_off;       // first tier-1 work item
    uint32_t xoff;           // first extra slot row
    uint16_t W;              // words per plane
    uint16_t cdmax;          // largest cd that passes the CDR3 gates
    uint16_t tiles;
    uint16_t seg;            // A tiles per segment
    uint32_t mma;            // MMA_NONE, or (bases of both CDR3s << 23 | first int8 tile slot): the group's tier-1 runs on the tensor cores (k_sweep_mma)
};
static_assert(sizeof(GroupInfo) == 32, "GroupInfo is 32 bytes");
__host__ __device__ __forceinline__ uint32_t seg_of_tiles(uint32_t T, uint32_t seg_min = SEG_MIN) { uint32_t s = (T + SEG_QMAX - 1) / SEG_QMAX; return s < seg_min ? seg_min : s; }
__host__ __device__ __forceinline__ uint32_t seg_item_off(uint32_t T, uint32_t S, uint32_t q) { return q * T - S * (q * (q - 1) / 2); }

struct DevCounters {         // device-resident scalars (one struct, copied back once)
    unsigned long long n_valid, n_groups, n_work, c3_words, pairs_compared, pairs_candidate, n_buckets;
    unsigned long long n_cand, n_todo, n_generic, n_generic_q, n_accept, n_kept, pairs_tier2;
    unsigned long long n_accept_r1, n_generic_r1, n_todo_r1, n_rest;
    unsigned long long work_next;   // dynamic tile scheduler
    unsigned long long pack_bytes[32];      // V..J bytes read + t2 row bytes written by k_pack_t2 (needed entries only)
    unsigned long long inplace_bytes[32];   // V..J bytes staged by k_pack_t2 (spread over 32 slots to keep the atomics apart)
    unsigned long long t1_alg_bytes, t2_alg_bytes, t2_units, n_classes;
    unsigned int err;
    unsigned int bor_active;
    unsigned int bor_done;           // a k_bor_min pass found no edge between two components: the later rounds return at once
    unsigned int n_payload_slow;
    unsigned int n_pending;
    unsigned int n_asc;              // 2-bit input: entries with an ASCII chain (k_meta lists them for k_pack_t2)
    unsigned long long n_xrows;      // extra slot rows asked for by the group table
    unsigned long long n_items;      // tier-1 work items of this rank (after the cost sort)
    unsigned int cls_count[64], cls_cursor[64];   // counting sort of the work items by cost class
    unsigned long long work_next_mma;               // (unused since k_sweep_mma deals the item list round-robin to its CTAs; kept for the struct layout)
    unsigned int i8_tiles, pad2;                    // int8 tile slots handed to tensor-core groups (k_group_build)
    unsigned long long mma_pairs, mma_macs;         // pairs of the tensor-core groups; int8 multiply-accumulates their tile pairs stand for
};

struct DevParams {           // ejoin_params_t plus run constants, by value to kernels
    double max_score, join_cdr3_ident, fwr1_cdr12_delta;
    int auto_share, comp_filt, comp_filt_bound, max_cdr3_diffs, cdr3_mult, max_degradation;
    int max_diffs, max_diffs_tcr, ref_v_trim, ref_j_trim, old_light, mix_donors, is_bcr, easy, degradation_mode;
};

// The entries a handle holds: up to EJOIN_MAX_RANGES ascending ranges of info indices, concatenated.  local index i of
// range r (base[r] <= i < base[r+1]) is info index lo[r] + (i - base[r]).
#define EJOIN_MAX_RANGES 16
struct RangeTab { uint32_t n; uint32_t lo[EJOIN_MAX_RANGES], base[EJOIN_MAX_RANGES]; };
__host__ __device__ __forceinline__ uint32_t rt_to_global(const RangeTab &rt, uint32_t local) {
    uint32_t r = 0;
    while (r + 1 < rt.n && rt.base[r + 1] <= local) r++;
    return rt.lo[r] + (local - rt.base[r]);
}
// info index -> local index (the index must belong to one of the ranges; n = total entries held)
__host__ __device__ __forceinline__ uint32_t rt_to_local(const RangeTab &rt, uint32_t n, uint32_t g) {
    for (uint32_t r = 0; r < rt.n; r++) {
        const uint32_t cnt = (r + 1 < rt.n ? rt.base[r + 1] : n) - rt.base[r];
        if (g >= rt.lo[r] && g - rt.lo[r] < cnt) return rt.base[r] + (g - rt.lo[r]);
    }
    return 0;
}

struct MultInfo { uint32_t off; uint32_t d1p1; };   // per (cl0,cl1): pool offset, row stride (D1+1); off = ~0 if absent

// Raw input arrays as uploaded (same meaning as ejoin_pack_t)
struct DevIn {
    uint32_t n_info, n_exact, n_refseq;
    const uint32_t *exact_id;
    const uint16_t *len[2];
    const uint64_t *tig_off[2];
    const uint8_t *tig_arena;       // V..J bytes as the scoring kernels read them (device memory)
    const uint8_t *tig_src;         // where k_pack_t2 reads them: == tig_arena, or the caller's pinned arena (zero-copy)
    uint8_t *tig_copy;              // zero-copy only: k_pack_t2 mirrors the needed entries' bytes here (== tig_arena)
    // EJOIN_F_TIGS_2BIT: chains without a deletion marker come 2 bits per base (CloneInfo.tigsp layout); tig_off of such a
    // chain is a word offset into tig2_src (device copy, or the caller's pinned arena read in place); tig_arena / tig_off
    // keep their ASCII meaning for the chains with has_del != 0.  nullptr = every chain is ASCII.
    const uint64_t *tig2_src;
    uint32_t tig2_planes;           // EJOIN_F_TIGS_PLANES: the words already hold the two bit planes
    // EJOIN_F_CDR3_2BIT: both CDR3s of an entry concatenated, 2 bits per base; nullptr = ASCII strings in cdr3_arena
    const uint64_t *cdr3_2bit;
    const uint32_t *cdr3_2bit_off;
    // uploaded slices (caller coordinates): ASCII V..J bytes, 2-bit V..J words, CDR3 bytes or words
    unsigned long long tig_lo, tig_hi, tig2_lo, tig2_hi, cdr3_lo, cdr3_hi;
    const uint8_t *has_del[2];
    const uint16_t *cdr3_len[2];
    const uint64_t *cdr3_off[2];
    const uint8_t *cdr3_arena;
    const uint16_t *cdr3_start[2];
    const uint32_t *v_ref_id[2], *j_ref_id[2], *dref_id[2], *vs_seq[2], *js_seq[2], *vname_id[2], *vuni_seq[2], *utr_seq[2];
    const uint32_t *c_ref_id_light;
    const uint16_t *hcomp, *fr1, *cdr1, *fr2, *cdr2, *fr3;
    const uint32_t *nchains, *ncells, *donor_lo, *donor_hi;
    const uint64_t *refseq_off;
    const uint32_t *refseq_len;
    const uint8_t *refseq_arena;
    // germline bit planes (k_pack_ref): per sequence ceil(len/32) words + one zero word before and after;
    // plane p of word w of sequence i = ref_planes[p * ref_plane_words + ref_plane_off[i] + w], p = H, L, N (not ACGT)
    const uint32_t *ref_plane_off;
    uint32_t *ref_planes;
    uint32_t ref_plane_words;
};

// Everything tier 2 needs
struct DevTables {
    const EntryMeta *meta;
    const uint32_t *c3;
    const uint4 *t2;
    const double *const *p1tab;     // [len0+len1] -> table [P1_DCAP][P1_KCAP] or null
    const MultInfo *multinfo;       // [65536]
    const double *multpool;
    const uint16_t *cdmax;          // [C3_MAX_TOTAL+1]
};

__device__ __forceinline__ int base_code(uint8_t b) {
    // A,C,G,T -> 0,1,2,3 ; anything else -> -1
    switch (b) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

// bits of 32-base word `w` (positions 32w..32w+31) that fall inside [lo,hi)
__device__ __forceinline__ uint32_t range_mask(int w, int lo, int hi) {
    int a = lo - 32 * w, b = hi - 32 * w;
    a = a < 0 ? 0 : a;
    b = b > 32 ? 32 : b;
    if (b <= a) return 0u;
    uint32_t m = (b >= 32) ? 0xFFFFFFFFu : ((1u << b) - 1u);
    return m & ~((a >= 32) ? 0xFFFFFFFFu : ((1u << a) - 1u));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk TMA copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
