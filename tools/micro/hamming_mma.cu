// hamming_mma.cu — PROTOTYPE (not on the product path): the tier-1 all-pairs CDR3 Hamming test of one giant bucket on the
// 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM), next to the XOR/popcount formulation the library
// ships (k_sweep).  VERDICT r01 "next" item 9; DESIGN.md §13 has the numbers and what integrating it would take.
//
// The contraction.  A base g in {A,C,G,T} becomes a vertex of the regular tetrahedron, v_g in {+1,-1}^3:
//   A (1,1,1)   C (1,-1,-1)   G (-1,1,-1)   T (-1,-1,1)          v_g . v_h = 3 if g == h, -1 otherwise.
// A CDR3 row of L bases is K = 3L int8 values (padded with zeros to a multiple of 32), and for two rows of one comparison
// group (equal L):   dot = 3 (L - d) - d = 3L - 4d,   so   d <= cdmax  <=>  dot >= 3L - 4 cdmax.      Exact in s32.
//
// The kernel.  One CTA (128 threads) keeps a B tile (128 entries = the M rows of the MMA) resident in shared memory and sweeps
// the A tiles 0..ib (128 entries = the N columns), as k_sweep does: 1-D bulk TMA copies (cp.async.bulk) of tiles that the pack
// kernel already wrote in the shared-memory image the MMA wants (K-major, no swizzle: [K/16][128 rows][16 bytes], core
// matrices of 8 rows x 16 bytes; descriptor LBO = 2048 B between the two 16-byte K chunks of one instruction, SBO = 128 B
// between groups of 8 rows), K/32 tcgen05.mma per tile pair issued by one thread, completion by tcgen05.commit on an
// mbarrier, two TMEM accumulator buffers of 128 columns so the MMA of tile i+1 runs under the epilogue of tile i; the epilogue
// is tcgen05.ld 32x32b.x32: thread t holds row b = t and 32 consecutive a's, exactly the (b, block of 32 a's) -> 32-bit
// survivor mask the library's survivor list is made of.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o hamming_mma hamming_mma.cu
// run:   ./hamming_mma [n=200000] [L=84] [cdmax=12] [reps=5]   -> one JSON line; exit 1 if the survivor sets differ
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(2); } } while (0)

typedef unsigned long long u64;
constexpr uint32_t TILE = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ u64 smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (u64)((addr >> 4) & 0x3FFF) | ((u64)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((u64)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, u64 da, u64 db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}

struct Tally { u64 pairs, sum_a, sum_b; };

// bases: [n][L] values 0..3.  tiles: [n_tiles][K/16][128][16] int8 (the shared-memory image of one operand tile)
__global__ void k_pack_i8(const uint8_t *bases, uint32_t n, uint32_t L, uint32_t K, int8_t *tiles, int extra) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;       // one thread per (entry slot, 16-byte chunk)
    const uint32_t chunks = K / 16;
    const size_t slots = (size_t)((n + TILE - 1) / TILE) * TILE;
    if (i >= slots * chunks) return;
    const uint32_t e = (uint32_t)(i / chunks), j = (uint32_t)(i % chunks);
    int8_t v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const uint32_t k = j * 16 + q, p = k / 3, c = k % 3;
        int8_t x = 0;
        if (e < n && p < L) {
            const uint32_t g = bases[(size_t)e * L + p];
            // A (1,1,1)  C (1,-1,-1)  G (-1,1,-1)  T (-1,-1,1)
            const bool neg = (g == 1 && c != 0) || (g == 2 && c != 1) || (g == 3 && c != 2);
            x = neg ? -1 : 1;
        }
        if (extra && k == 3 * L) x = 16;            // version 4: the a side of the two threshold slots (the b side is patched in shared memory)
        if (extra && k == 3 * L + 1) x = 1;
        v[q] = x;
    }
    int8_t *dst = tiles + (size_t)(e / TILE) * TILE * K + (size_t)j * TILE * 16 + (size_t)(e % TILE) * 16;
    *reinterpret_cast<int4 *>(dst) = *reinterpret_cast<const int4 *>(v);
}

// 2-bit planes for the reference: [n][2][W] words (H plane, L plane), as the library's c3 rows
__global__ void k_pack_planes(const uint8_t *bases, uint32_t n, uint32_t L, uint32_t W, uint32_t *rows) {
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    for (uint32_t w = 0; w < W; w++) {
        uint32_t h = 0, l = 0;
        for (uint32_t q = 0; q < 32; q++) {
            const uint32_t p = w * 32 + q;
            if (p < L) { const uint32_t g = bases[e * L + p]; h |= (g >> 1) << q; l |= (g & 1u) << q; }
        }
        rows[e * 2 * W + w] = h; rows[e * 2 * W + W + w] = l;
    }
}
// reference: the popcount formulation, one thread per b, plain loop (a checker, not the library's tiled k_sweep)
__global__ void k_ref_popc(const uint32_t *rows, uint32_t n, uint32_t W, uint32_t cdmax, Tally *out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    u64 pairs = 0, sa = 0, sb = 0;
    if (b < n) {
        uint32_t bh[4], bl[4];
        for (uint32_t w = 0; w < W; w++) { bh[w] = rows[(size_t)b * 2 * W + w]; bl[w] = rows[(size_t)b * 2 * W + W + w]; }
        for (uint32_t a = 0; a < b; a++) {
            uint32_t d = 0;
            for (uint32_t w = 0; w < W; w++) d += __popc((bh[w] ^ rows[(size_t)a * 2 * W + w]) | (bl[w] ^ rows[(size_t)a * 2 * W + W + w]));
            if (d <= cdmax) { pairs++; sa += a; sb += b; }
        }
    }
    for (int o = 16; o; o >>= 1) { pairs += __shfl_down_sync(~0u, pairs, o); sa += __shfl_down_sync(~0u, sa, o); sb += __shfl_down_sync(~0u, sb, o); }
    if ((threadIdx.x & 31) == 0 && pairs) { atomicAdd(&out->pairs, pairs); atomicAdd(&out->sum_a, sa); atomicAdd(&out->sum_b, sb); }
}

// The tensor-core sweep.  dynamic smem: sB [128 x K] | sA[0] | sA[1]
// dbg: bit 0 = every A tile load fetches tile 0 again (same bytes, always L2 hits on one line set); bit 1 = skip the compares (TMEM loads kept): timing experiments only
__global__ void __launch_bounds__(128) k_hamming_mma(const int8_t *__restrict__ tiles, uint32_t n, uint32_t K, int thr, u64 *work_next, Tally *out, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_done[2], bar_b;
    __shared__ uint32_t s_tmem, s_item;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t tile_bytes = TILE * K;
    uint8_t *sB = smem, *sA0 = smem + tile_bytes, *sA1 = smem + 2 * tile_bytes;
    const uint32_t n_tiles = (n + TILE - 1) / TILE;
    if (tid == 0) { mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1); mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1); mbar_init(&bar_b, 1); fence_barrier_init(); }
    if (warp == 0) {        // 2 accumulator buffers x 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): D = s32, A = B = signed 8 bit, both K-major, N = 128, M = 128
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t ksteps = K / 32;
    uint32_t ph_full = 0, ph_done = 0, ph_b = 0;        // bit s = phase of the barrier of buffer s
    u64 pairs = 0, sum_a = 0, sum_b = 0;

    auto issue_mma = [&](uint32_t s) {                   // thread 0: the K/32 instructions of one tile pair into accumulator buffer s
        const uint32_t a_addr = smem_u32(sB), b_addr = smem_u32(s ? sA1 : sA0);
        for (uint32_t k = 0; k < ksteps; k++) {
            const u64 da = smem_desc(a_addr + k * 2 * TILE * 16, TILE * 16, 128);
            const u64 db = smem_desc(b_addr + k * 2 * TILE * 16, TILE * 16, 128);
            mma_i8(tmem + s * 128, da, db, idesc, k ? 1u : 0u);
        }
        mma_commit(&bar_done[s]);
    };
    auto load_a = [&](uint32_t ia, uint32_t s) {
        mbar_expect_tx(&bar_full[s], tile_bytes);
        tma_load_1d(s ? sA1 : sA0, tiles + (size_t)((dbg & 1) ? 0 : ia) * tile_bytes, tile_bytes, &bar_full[s]);
    };

    for (;;) {
        if (tid == 0) s_item = (uint32_t)atomicAdd(work_next, 1ull);
        __syncthreads();
        const uint32_t item = s_item;
        __syncthreads();
        if (item >= n_tiles) break;
        const uint32_t ib = n_tiles - 1 - item;          // longest rows first
        const uint32_t b = ib * TILE + tid;
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bar_b, tile_bytes);
            tma_load_1d(sB, tiles + (size_t)ib * tile_bytes, tile_bytes, &bar_b);
            load_a(0, 0);
            if (ib >= 1) load_a(1, 1);
            mbar_wait(&bar_b, ph_b & 1u);
            mbar_wait(&bar_full[0], ph_full & 1u); ph_full ^= 1u;
            tc_fence_after();
            issue_mma(0);
        }
        ph_b ^= 1u;
        for (uint32_t ia = 0; ia <= ib; ia++) {
            const uint32_t s = ia & 1u;
            mbar_wait(&bar_done[s], (ph_done >> s) & 1u);            // MMA(ia) complete: accumulator s is ready, smem buffer s is free
            ph_done ^= 1u << s;
            tc_fence_after();
            if (tid == 0) {
                if (ia + 2 <= ib) { fence_proxy_async(); load_a(ia + 2, s); }
                if (ia + 1 <= ib) {
                    mbar_wait(&bar_full[s ^ 1u], (ph_full >> (s ^ 1u)) & 1u); ph_full ^= 1u << (s ^ 1u);
                    tc_fence_after();
                    issue_mma(s ^ 1u);                                // runs under the epilogue below
                }
            }
            __syncwarp();
            // epilogue: row b = tid (TMEM lane), columns = the 128 a's of tile ia
            const bool diag = ia == ib;
            // all the TMEM loads of the row first (four x32 loads in flight, one wait), then the compares
            uint32_t r[4][32];
            const uint32_t nblk = diag ? min(4u, ((tid | 31u) >> 5) + 1u) : 4u;      // warp-uniform: blocks of 32 a's that can hold an a < b
            const uint32_t trow = tmem + ((warp * 32u) << 16) + s * 128;
            tmem_ld32_nowait(trow, r[0]);
            if (nblk > 1) tmem_ld32_nowait(trow + 32, r[1]);
            if (nblk > 2) tmem_ld32_nowait(trow + 64, r[2]);
            if (nblk > 3) tmem_ld32_nowait(trow + 96, r[3]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (uint32_t blk = 0; blk < 4; blk++) {
                if (blk >= nblk) break;
                if (dbg & 2) { if (r[blk][0] == 0x7FFFFFFF) pairs++; continue; }
                const uint32_t c0 = blk * 32;
                uint32_t bits = 0;
                // bit q = (dot >= thr): the sign bit of (thr - 1 - dot), shifted in by one funnel shift per element
                // (four independent chains of eight, so that one warp per scheduler is not waiting on a 32-deep dependency chain)
                uint32_t part[4] = { 0, 0, 0, 0 };
#pragma unroll
                for (int q = 7; q >= 0; q--)
#pragma unroll
                    for (int c = 0; c < 4; c++) part[c] = __funnelshift_l((uint32_t)(thr - 1 - (int)r[blk][8 * c + q]), part[c], 1);
                bits = part[0] | (part[1] << 8) | (part[2] << 16) | (part[3] << 24);
                const uint32_t a0 = ia * TILE + c0;
                if (b >= n) bits = 0;
                if (a0 + 32 > n) bits &= a0 < n ? (0xFFFFFFFFu >> (a0 + 32 - n)) : 0u;
                if (diag) {                                           // keep a < b
                    if (tid <= c0) bits = 0;
                    else if (tid - c0 < 32) bits &= (1u << (tid - c0)) - 1u;
                }
                while (bits) { const uint32_t q = __ffs(bits) - 1; bits &= bits - 1; pairs++; sum_a += a0 + q; sum_b += b; }
            }
            tc_fence_before();
            __syncthreads();                                          // accumulator s drained before MMA(ia + 2) may overwrite it
        }
    }
    for (int o = 16; o; o >>= 1) { pairs += __shfl_down_sync(~0u, pairs, o); sum_a += __shfl_down_sync(~0u, sum_a, o); sum_b += __shfl_down_sync(~0u, sum_b, o); }
    if ((tid & 31) == 0 && pairs) { atomicAdd(&out->pairs, pairs); atomicAdd(&out->sum_a, sum_a); atomicAdd(&out->sum_b, sum_b); }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// Version 2: warp-specialised.  Warps 0-3 = epilogue (thread = row b), warp 4 lane 0 = MMA issuer, warp 5 lane 0 = TMA producer; NA
// shared-memory stages for the A tiles and NACC accumulator stages in TMEM, so the issuer queues the MMAs of the next tiles
// behind the running one instead of waiting for it (version 1 waits for MMA(i) before it issues MMA(i+1)), and the epilogue
// frees an accumulator as soon as it has read it.  Items (B tiles) are dealt round-robin, longest first; every role walks the
// same item sequence with its own stage counter.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// SIGN (version 4): the threshold rides in the contraction — two extra K slots hold (16, 1) on the a side and (-thr / 16, -thr % 16) on the
// b side (patched into the B tile in shared memory by the epilogue threads), so the accumulator is dot - thr and a survivor is a clear sign bit:
// one funnel shift per element instead of a subtraction and a shift.
template <uint32_t NA, uint32_t NACC, bool SIGN>
__global__ void __launch_bounds__(192) k_hamming_mma2(const int8_t *__restrict__ tiles, uint32_t n, uint32_t K, int thr, Tally *out, uint32_t L) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[NA], bar_sfree[NA], bar_done[NACC], bar_empty[NACC], bar_bfull, bar_bfree, bar_bpatch;
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tile_bytes = TILE * K;
    uint8_t *sB = smem;
    const uint32_t n_tiles = (n + TILE - 1) / TILE;
    if (tid == 0) {
        for (uint32_t i = 0; i < NA; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_sfree[i], 1); }
        for (uint32_t i = 0; i < NACC; i++) { mbar_init(&bar_done[i], 1); mbar_init(&bar_empty[i], 4); }
        mbar_init(&bar_bfull, 1); mbar_init(&bar_bfree, 1); mbar_init(&bar_bpatch, 4);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(NACC * 128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t ksteps = K / 32;
    if (warp == 5) {                                     // ---- TMA producer ----
        if (lane == 0) {
            uint32_t t = 0, item_no = 0;
            for (uint32_t item = blockIdx.x; item < n_tiles; item += gridDim.x, item_no++) {
                const uint32_t ib = n_tiles - 1 - item;
                mbar_wait(&bar_bfree, (item_no & 1u) ^ 1u);                 // the MMAs of the previous item are done with sB
                fence_proxy_async();
                mbar_expect_tx(&bar_bfull, tile_bytes);
                tma_load_1d(sB, tiles + (size_t)ib * tile_bytes, tile_bytes, &bar_bfull);
                for (uint32_t ia = 0; ia <= ib; ia++, t++) {
                    const uint32_t st = t % NA;
                    mbar_wait(&bar_sfree[st], ((t / NA) & 1u) ^ 1u);
                    mbar_expect_tx(&bar_full[st], tile_bytes);
                    tma_load_1d(smem + (size_t)(1 + st) * tile_bytes, tiles + (size_t)ia * tile_bytes, tile_bytes, &bar_full[st]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 4) {                              // ---- MMA issuer ----
        if (lane == 0) {
            uint32_t t = 0, item_no = 0;
            for (uint32_t item = blockIdx.x; item < n_tiles; item += gridDim.x, item_no++) {
                const uint32_t ib = n_tiles - 1 - item;
                mbar_wait(SIGN ? &bar_bpatch : &bar_bfull, item_no & 1u);
                for (uint32_t ia = 0; ia <= ib; ia++, t++) {
                    const uint32_t st = t % NA, ac = t % NACC;
                    mbar_wait(&bar_full[st], (t / NA) & 1u);
                    mbar_wait(&bar_empty[ac], ((t / NACC) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sB), b_addr = smem_u32(smem + (size_t)(1 + st) * tile_bytes);
                    for (uint32_t k = 0; k < ksteps; k++)
                        mma_i8(tmem + ac * 128, smem_desc(a_addr + k * 2 * TILE * 16, TILE * 16, 128), smem_desc(b_addr + k * 2 * TILE * 16, TILE * 16, 128), idesc, k ? 1u : 0u);
                    mma_commit(&bar_sfree[st]);
                    mma_commit(&bar_done[ac]);
                }
                mma_commit(&bar_bfree);
            }
        }
        __syncwarp();
    } else {                                             // ---- epilogue: thread = row b ----
        u64 pairs = 0, sum_a = 0, sum_b = 0;
        uint32_t t = 0, item_no = 0;
        for (uint32_t item = blockIdx.x; item < n_tiles; item += gridDim.x, item_no++) {
            const uint32_t ib = n_tiles - 1 - item;
            const uint32_t b = ib * TILE + tid;
            if (SIGN) {                                  // the b side of the threshold slots, into this thread's row of the B tile
                mbar_wait(&bar_bfull, item_no & 1u);
                const uint32_t k0 = 3 * L, k1 = 3 * L + 1;
                sB[(k0 / 16) * TILE * 16 + tid * 16 + (k0 % 16)] = (uint8_t)(int8_t)(-(thr / 16));
                sB[(k1 / 16) * TILE * 16 + tid * 16 + (k1 % 16)] = (uint8_t)(int8_t)(-(thr % 16));
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_bpatch);
            }
            for (uint32_t ia = 0; ia <= ib; ia++, t++) {
                const uint32_t ac = t % NACC;
                mbar_wait(&bar_done[ac], (t / NACC) & 1u);
                tc_fence_after();
                const bool diag = ia == ib;
                uint32_t r[4][32];
                const uint32_t nblk = diag ? min(4u, ((tid | 31u) >> 5) + 1u) : 4u;
                const uint32_t trow = tmem + ((warp * 32u) << 16) + ac * 128;
                tmem_ld32_nowait(trow, r[0]);
                if (nblk > 1) tmem_ld32_nowait(trow + 32, r[1]);
                if (nblk > 2) tmem_ld32_nowait(trow + 64, r[2]);
                if (nblk > 3) tmem_ld32_nowait(trow + 96, r[3]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[ac]);                 // this warp has its 32 rows of the accumulator in registers
#pragma unroll
                for (uint32_t blk = 0; blk < 4; blk++) {
                    if (blk >= nblk) break;
                    const uint32_t c0 = blk * 32;
                    uint32_t part[4] = { 0, 0, 0, 0 };
#pragma unroll
                    for (int q = 7; q >= 0; q--)
#pragma unroll
                        for (int c = 0; c < 4; c++) part[c] = __funnelshift_l(SIGN ? r[blk][8 * c + q] : (uint32_t)(thr - 1 - (int)r[blk][8 * c + q]), part[c], 1);
                    uint32_t bits = part[0] | (part[1] << 8) | (part[2] << 16) | (part[3] << 24);
                    if (SIGN) bits = ~bits;                  // a survivor is a clear sign bit of dot - thr
                    const uint32_t a0 = ia * TILE + c0;
                    if (b >= n) bits = 0;
                    if (a0 + 32 > n) bits &= a0 < n ? (0xFFFFFFFFu >> (a0 + 32 - n)) : 0u;
                    if (diag) {
                        if (tid <= c0) bits = 0;
                        else if (tid - c0 < 32) bits &= (1u << (tid - c0)) - 1u;
                    }
                    while (bits) { const uint32_t q = __ffs(bits) - 1; bits &= bits - 1; pairs++; sum_a += a0 + q; sum_b += b; }
                }
            }
        }
        for (int o = 16; o; o >>= 1) { pairs += __shfl_down_sync(~0u, pairs, o); sum_a += __shfl_down_sync(~0u, sum_a, o); sum_b += __shfl_down_sync(~0u, sum_b, o); }
        if (lane == 0 && pairs) { atomicAdd(&out->pairs, pairs); atomicAdd(&out->sum_a, sum_a); atomicAdd(&out->sum_b, sum_b); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(NACC * 128u) : "memory");
}

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 200000u;
    const uint32_t L = argc > 2 ? (uint32_t)atoi(argv[2]) : 84u;
    const uint32_t cdmax = argc > 3 ? (uint32_t)atoi(argv[3]) : 12u;
    const int reps = argc > 4 ? atoi(argv[4]) : 5;
    const int dbg = argc > 6 ? atoi(argv[6]) : 0;
    const int version = argc > 7 ? atoi(argv[7]) : 2;
    if (L == 0 || L > 128 || n == 0) { fprintf(stderr, "1 <= L <= 128\n"); return 2; }
    const uint32_t K = (3 * L + 2 + 31) / 32 * 32, W = (L + 31) / 32;      // room for the two threshold slots of version 4
    const uint32_t n_tiles = (n + TILE - 1) / TILE;
    // a giant bucket: families of near-identical CDR3s (a few substitutions from a founder) among unrelated ones
    std::vector<uint8_t> h((size_t)n * L);
    u64 rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    for (uint32_t e = 0; e < n; e++) {
        const bool fam = e > 0 && (next() % 100) < 60;
        const uint32_t src = fam ? (uint32_t)(next() % e) : e;
        for (uint32_t p = 0; p < L; p++) h[(size_t)e * L + p] = fam ? h[(size_t)src * L + p] : (uint8_t)(next() & 3);
        if (fam) for (uint32_t m = next() % 9; m; m--) h[(size_t)e * L + next() % L] = (uint8_t)(next() & 3);
    }
    uint8_t *d_bases; int8_t *d_tiles; uint32_t *d_rows; Tally *d_t; u64 *d_work;
    CK(cudaMalloc(&d_bases, h.size())); CK(cudaMemcpy(d_bases, h.data(), h.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_tiles, (size_t)n_tiles * TILE * K)); CK(cudaMalloc(&d_rows, (size_t)n * 2 * W * 4)); CK(cudaMalloc(&d_t, 2 * sizeof(Tally))); CK(cudaMalloc(&d_work, 8));
    CK(cudaMemset(d_t, 0, 2 * sizeof(Tally)));
    const size_t pack_threads = (size_t)n_tiles * TILE * (K / 16);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_pack_i8<<<(unsigned)((pack_threads + 255) / 256), 256>>>(d_bases, n, L, K, d_tiles, version == 4);
    CK(cudaEventRecord(e1));
    k_pack_planes<<<(n + 255) / 256, 256>>>(d_bases, n, L, W, d_rows);
    CK(cudaDeviceSynchronize());
    float ms_pack = 0; cudaEventElapsedTime(&ms_pack, e0, e1);
    const size_t smem = 3 * (size_t)TILE * K, smem2 = 4 * (size_t)TILE * K;
    CK(cudaFuncSetAttribute(k_hamming_mma2<3, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    CK(cudaFuncSetAttribute(k_hamming_mma2<2, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_hamming_mma2<2, 2, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_hamming_mma2<2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_hamming_mma2<2, 2, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_hamming_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_hamming_mma, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int dev = 0, sms = 0, occ = 0;
    CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_hamming_mma, 128, smem));
    if (occ < 1) { fprintf(stderr, "kernel does not fit\n"); return 2; }
    const int occ_api = occ;
    occ = argc > 5 ? atoi(argv[5]) : 2;            // CTAs per SM launched: 256 TMEM columns per CTA, 512 per SM
    fprintf(stderr, "occupancy API: %d CTAs per SM; launching %d per SM\n", occ_api, occ);
    const int thr = (int)(3 * L) - 4 * (int)cdmax;
    float best = 1e30f;
    Tally got{};
    for (int r = 0; r < reps + 1; r++) {
        CK(cudaMemset(d_work, 0, 8)); CK(cudaMemset(d_t, 0, sizeof(Tally)));
        CK(cudaEventRecord(e0));
        if (version == 2) k_hamming_mma2<3, 4, false><<<sms, 192, smem2>>>(d_tiles, n, K, thr, d_t, L);
        else if (version == 3) k_hamming_mma2<2, 2, false><<<sms * 2, 192, smem>>>(d_tiles, n, K, thr, d_t, L);
        else if (version == 4) k_hamming_mma2<2, 2, true><<<sms * 2, 192, smem>>>(d_tiles, n, K, thr, d_t, L);
        else k_hamming_mma<<<sms * occ, 128, smem>>>(d_tiles, n, K, thr, d_work, d_t, dbg);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
        CK(cudaMemcpy(&got, d_t, sizeof(Tally), cudaMemcpyDeviceToHost));
    }
    CK(cudaEventRecord(e0));
    k_ref_popc<<<(n + 127) / 128, 128>>>(d_rows, n, W, cdmax, d_t + 1);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms_ref = 0; cudaEventElapsedTime(&ms_ref, e0, e1);
    Tally want{};
    CK(cudaMemcpy(&want, d_t + 1, sizeof(Tally), cudaMemcpyDeviceToHost));
    const bool same = got.pairs == want.pairs && got.sum_a == want.sum_a && got.sum_b == want.sum_b;
    const double pair_count = (double)n * (n - 1) / 2, macs = (double)n_tiles * (n_tiles + 1) / 2 * TILE * TILE * K;
    printf("{\"tool\": \"hamming_mma\", \"n\": %u, \"L\": %u, \"K\": %u, \"cdmax\": %u, \"ctas_per_sm\": %d, \"ms_mma_sweep\": %.3f, \"ms_pack_i8\": %.3f, \"pairs_per_s\": %.4g, "
           "\"int8_tmacs\": %.1f, \"survivors\": %llu, \"survivors_ref\": %llu, \"same_survivor_set\": %s, \"ms_ref_plain_popc_kernel\": %.3f, \"dbg\": %d, \"version\": %d}\n",
           n, L, K, cdmax, occ, best, ms_pack, pair_count / (best * 1e-3), macs / (best * 1e-3) / 1e12, got.pairs, want.pairs, same ? "true" : "false", ms_ref, dbg, version);
    if (dbg) return 0;
    return same ? 0 : 1;
}
