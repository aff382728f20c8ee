// fp64 micro-benchmarks on sm_100a: DFMA/DADD/DMUL latency and throughput, divide and sqrt
// latency/throughput, exp cost, Thomas-chain cycles per level.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void lat_fma(double *out, double a, double b, int n, long long *cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) x = __fma_rn(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lat_add(double *out, double a, int n, long long *cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) x = __dadd_rn(x, a);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lat_div(double *out, double a, int n, long long *cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n; i++) x = a / x;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lat_sqrt(double *out, double a, int n, long long *cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n; i++) x = sqrt(x + a);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// throughput: many independent chains per thread, many warps
template <int ILP>
__global__ void thr_fma(double *out, double a, double b, int n)
{
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = out[(blockIdx.x * blockDim.x + threadIdx.x) * ILP + j];
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = __fma_rn(x[j], a, b);
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) out[(blockIdx.x * blockDim.x + threadIdx.x) * ILP + j] = x[j];
}
template <int ILP>
__global__ void thr_div(double *out, double a, int n)
{
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = out[(blockIdx.x * blockDim.x + threadIdx.x) * ILP + j];
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = a / x[j];
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) out[(blockIdx.x * blockDim.x + threadIdx.x) * ILP + j] = x[j];
}
// Thomas forward chain: per level  gam = cl/bet; bet = cc - cu*gam; yn = (rhs - cu*yn)/bet   (coefficients in smem)
__global__ void thomas_chain(double *out, const double *coef, int nlev, int reps, long long *cyc)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4 * nlev; i += blockDim.x) sm[i] = coef[i];
    __syncthreads();
    double bet = sm[0], yn = sm[3 * nlev] / bet, acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        for (int i = 1; i < nlev; i++) {
            const double cu = sm[i], cc = sm[nlev + i], cl = sm[2 * nlev + i - 1], rhs = sm[3 * nlev + i];
            const double gam = cl / bet;
            bet = cc - cu * gam;
            yn = (rhs - cu * yn) / bet;
        }
        acc += yn;
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc + bet;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *d; long long *c;
    cudaMalloc(&d, 1 << 26); cudaMemset(d, 0, 1 << 26);
    cudaMallocManaged(&c, 8);
    std::vector<double> h(1 << 16, 1.25);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    const int n = 4096;
    lat_fma<<<1, 32>>>(d, 0.999, 0.001, n, c); cudaDeviceSynchronize(); printf("DFMA dependent latency: %.2f cycles\n", (double)*c / n);
    lat_add<<<1, 32>>>(d, 0.001, n, c); cudaDeviceSynchronize(); printf("DADD dependent latency: %.2f cycles\n", (double)*c / n);
    lat_div<<<1, 32>>>(d, 1.7, n, c); cudaDeviceSynchronize(); printf("DDIV dependent latency: %.2f cycles\n", (double)*c / n);
    lat_sqrt<<<1, 32>>>(d, 1.7, n, c); cudaDeviceSynchronize(); printf("DSQRT(+add) dependent latency: %.2f cycles\n", (double)*c / n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    {
        const int blocks = 148 * 8, threads = 256, iters = 20000;
        thr_fma<8><<<blocks, threads>>>(d, 0.999, 0.001, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); thr_fma<8><<<blocks, threads>>>(d, 0.999, 0.001, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * threads * 8 * iters;
        printf("DFMA throughput: %.2f T DFMA/s = %.2f TFLOP/s  (%.1f DFMA/clk/SM at 1.965 GHz)\n", ops / ms / 1e9, 2 * ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
    }
    {
        const int blocks = 148 * 8, threads = 256, iters = 2000;
        thr_div<4><<<blocks, threads>>>(d, 1.7, 10); cudaDeviceSynchronize();
        cudaEventRecord(e0); thr_div<4><<<blocks, threads>>>(d, 1.7, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * threads * 4 * iters;
        printf("DDIV throughput: %.3f T div/s  (%.2f div/clk/SM) => one divide = %.1f DFMA slots\n", ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9, 64.0 / (ops / (ms * 1e-3) / 148 / 1.965e9));
    }
    {
        const int nlev = 100;
        std::vector<double> coef(4 * nlev);
        for (int i = 0; i < nlev; i++) { coef[i] = -0.3; coef[nlev + i] = 1.7; coef[2 * nlev + i] = -0.4; coef[3 * nlev + i] = 1.0 + 0.01 * i; }
        double *dc; cudaMalloc(&dc, coef.size() * 8); cudaMemcpy(dc, coef.data(), coef.size() * 8, cudaMemcpyHostToDevice);
        thomas_chain<<<1, 32, 4 * nlev * 8>>>(d, dc, nlev, 50, c); cudaDeviceSynchronize();
        printf("Thomas forward chain (1 warp, coefficients in smem): %.1f cycles per level\n", (double)*c / (50.0 * (nlev - 1)));
        thomas_chain<<<1, 128, 4 * nlev * 8>>>(d, dc, nlev, 50, c); cudaDeviceSynchronize();
        printf("Thomas forward chain (4 warps): %.1f cycles per level\n", (double)*c / (50.0 * (nlev - 1)));
    }
    return 0;
}
