// zc_bench.cu — how fast can SMs read pinned host memory in place (zero-copy), by access shape?  Sets the ceiling for the
// in-place pull of the needed entries' 2-bit V..J words (k_pack_t2), next to the copy engine's rate for the same bytes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zc_bench zc_bench.cu     run: ./zc_bench [MB]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <typename T> __global__ void rd_dense(const T *src, size_t n, unsigned long long *sink) {
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T v = src[i];
        const unsigned long long *w = (const unsigned long long *)&v;
        for (int q = 0; q < (int)(sizeof(T) / 8); q++) acc ^= w[q];
    }
    if (acc == 0x123456789ull) *sink = acc;
}
// entries of `ew` 8-byte words, every other entry skipped (the "needed" pattern): 16 lanes per entry, lane = word
__global__ void rd_entries(const unsigned long long *src, size_t n_entries, int ew, int keep_mod, unsigned long long *sink) {
    unsigned long long acc = 0;
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    for (size_t e = g; e < n_entries; e += ((size_t)gridDim.x * blockDim.x) >> 4) {
        if (keep_mod > 1 && (e * 2654435761u >> 7) % keep_mod) continue;
        for (int w = lane; w < ew; w += 16) acc ^= src[e * ew + w];
    }
    if (acc == 0x123456789ull) *sink = acc;
}

int main(int argc, char **argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 256;
    const size_t bytes = mb << 20;
    void *h; cudaMallocHost(&h, bytes);
    for (size_t i = 0; i < bytes / 8; i++) ((unsigned long long *)h)[i] = i * 0x9E3779B97F4A7C15ull;
    void *d; cudaMalloc(&d, bytes);
    unsigned long long *sink; cudaMalloc(&sink, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        if (rep) printf("copy engine H2D          : %7.2f GB/s\n", bytes / ms / 1e6);
    }
    const int grids[] = { 148 * 2, 148 * 8, 148 * 32 };
    for (int g : grids) {
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(a); rd_dense<unsigned long long><<<g, 256>>>((const unsigned long long *)h, bytes / 8, sink); cudaEventRecord(b); cudaEventSynchronize(b); }
        cudaEventElapsedTime(&ms, a, b); printf("SM read dense  8 B/lane, grid %5d: %7.2f GB/s\n", g, bytes / ms / 1e6);
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(a); rd_dense<uint4><<<g, 256>>>((const uint4 *)h, bytes / 16, sink); cudaEventRecord(b); cudaEventSynchronize(b); }
        cudaEventElapsedTime(&ms, a, b); printf("SM read dense 16 B/lane, grid %5d: %7.2f GB/s\n", g, bytes / ms / 1e6);
    }
    for (int ew : { 22, 24, 32 })
        for (int keep : { 1, 2 }) {
            const size_t ne = bytes / 8 / ew;
            for (int rep = 0; rep < 2; rep++) { cudaEventRecord(a); rd_entries<<<148 * 16, 256>>>((const unsigned long long *)h, ne, ew, keep, sink); cudaEventRecord(b); cudaEventSynchronize(b); }
            cudaEventElapsedTime(&ms, a, b);
            printf("SM read entries of %2d words, 1 of %d kept: %7.2f GB/s of kept bytes\n", ew, keep, (double)bytes / keep / ms / 1e6);
        }
    cudaError_t e = cudaGetLastError();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
