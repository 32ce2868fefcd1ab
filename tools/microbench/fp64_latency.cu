// Dependent-issue latency of the fp64 pipe and of an L2-hit load on sm_100a (one warp, clock64 around a chain).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_dfma(double *out, long long *cyc, double a, double b, int n) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int k = 0; k < n; ++k) x = __fma_rn(x, b, a);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_dadd(double *out, long long *cyc, double a, double b, int n) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int k = 0; k < n; ++k) x = __dadd_rn(x, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// four independent chains per thread (the four divisions of a node)
__global__ void chain_dfma4(double *out, long long *cyc, double a, double b, int n) {
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3;
    long long t0 = clock64();
#pragma unroll 8
    for (int k = 0; k < n; ++k) {
        x0 = __fma_rn(x0, b, a);
        x1 = __fma_rn(x1, b, a);
        x2 = __fma_rn(x2, b, a);
        x3 = __fma_rn(x3, b, a);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x0 + x1 + x2 + x3;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// pointer chase through L2 (ld.global.cg), 8 bytes per hop
__global__ void chase(const unsigned long long *p, long long *cyc, unsigned long long *sink, int n) {
    unsigned long long i = threadIdx.x;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) i = __ldcg(p + i);
    long long t1 = clock64();
    sink[threadIdx.x] = i;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 8 * 1024); cudaMallocManaged(&cyc, 8);
    const int n = 4096;
    for (int warps : {1, 8, 32}) {
        chain_dfma<<<1, 32 * warps>>>(out, cyc, 1.0, 0.999, n); cudaDeviceSynchronize();
        chain_dfma<<<1, 32 * warps>>>(out, cyc, 1.0, 0.999, n); cudaDeviceSynchronize();
        printf("DFMA dependent chain, %2d warps in one CTA: %.1f cycles per op\n", warps, (double)cyc[0] / n);
        chain_dadd<<<1, 32 * warps>>>(out, cyc, 1.0, 0.5, n); cudaDeviceSynchronize();
        printf("DADD dependent chain, %2d warps in one CTA: %.1f cycles per op\n", warps, (double)cyc[0] / n);
        chain_dfma4<<<1, 32 * warps>>>(out, cyc, 1.0, 0.999, n); cudaDeviceSynchronize();
        printf("4 independent DFMA chains, %2d warps: %.1f cycles per round of 4\n", warps, (double)cyc[0] / n);
    }
    const size_t N = 1 << 22;  // 32 MB: L2-resident, far beyond L1
    unsigned long long *h = (unsigned long long *)malloc(N * 8), *d, *sink;
    for (size_t i = 0; i < N; ++i) h[i] = (i * 4099 + 1031 * 64) % N;
    cudaMalloc(&d, N * 8); cudaMalloc(&sink, 8 * 1024);
    cudaMemcpy(d, h, N * 8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; ++rep) { chase<<<1, 32>>>(d, cyc, sink, 2000); cudaDeviceSynchronize(); }
    printf("ld.global.cg pointer chase (L2-resident 32 MB, one warp, divergent addresses): %.0f cycles per hop\n", (double)cyc[0] / 2000);
    for (int rep = 0; rep < 2; ++rep) { chase<<<1, 1>>>(d, cyc, sink, 2000); cudaDeviceSynchronize(); }
    printf("ld.global.cg pointer chase (one thread): %.0f cycles per hop\n", (double)cyc[0] / 2000);
    return 0;
}
