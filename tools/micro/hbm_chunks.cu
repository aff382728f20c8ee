// Achievable HBM bandwidth on B200 for chunked "random" access: every warp reads (and optionally
// writes) CHUNK contiguous bytes at a pseudo-random CHUNK-aligned offset of a buffer >> L2.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <int CHUNK, bool WRITE>
__global__ void k(double *buf, size_t nchunks, int iters, double *sink)
{
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    double acc = 0;
    constexpr int PER = CHUNK / 256;   // 256 B (32 lanes x 8 B) pieces per chunk
    for (int it = 0; it < iters; it++) {
        const size_t ch = mix(warp * 1315423911ull + it) % nchunks;
        double *p = buf + ch * (CHUNK / 8);
#pragma unroll
        for (int q = 0; q < PER; q++) {
            double v = p[q * 32 + lane];
            if (WRITE) p[q * 32 + lane] = v + 1.0; else acc += v;
        }
    }
    if (acc == 12345.678) sink[0] = acc;
}

template <int CHUNK, bool WRITE>
void run(double *buf, size_t bytes, double *sink)
{
    const size_t nchunks = bytes / CHUNK;
    const int iters = 4096 * 256 / CHUNK;
    const int blocks = 148 * 16, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CHUNK, WRITE><<<blocks, threads>>>(buf, nchunks, 8, sink);
    cudaEventRecord(e0);
    k<CHUNK, WRITE><<<blocks, threads>>>(buf, nchunks, iters, sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double moved = (double)blocks * threads / 32 * iters * CHUNK * (WRITE ? 2 : 1);
    printf("chunk %5d B %s: %.0f GB/s\n", CHUNK, WRITE ? "read+write" : "read      ", moved / ms / 1e6);
}

int main()
{
    const size_t bytes = 8ull << 30;
    double *buf, *sink; cudaMalloc(&buf, bytes); cudaMalloc(&sink, 8); cudaMemset(buf, 0, bytes);
    run<256, false>(buf, bytes, sink); run<512, false>(buf, bytes, sink); run<1024, false>(buf, bytes, sink);
    run<2048, false>(buf, bytes, sink); run<4096, false>(buf, bytes, sink);
    run<256, true>(buf, bytes, sink); run<1024, true>(buf, bytes, sink); run<4096, true>(buf, bytes, sink);
    return 0;
}
