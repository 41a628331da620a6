// clock_probe.cu -- sustained FP64 DMMA throughput and effective SM clock over a multi-second run.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, unsigned long long* clk, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    unsigned long long t0, t1; long long c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    long long c1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = t1 - t0; clk[1] = c1 - c0; }
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    double* d; unsigned long long* clk; cudaMalloc(&d, 64); cudaMallocManaged(&clk, 16);
    int blocks = pr.multiProcessorCount * 2, threads = 256, iters = 200000;
    for (int rep = 0; rep < 12; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<blocks, threads>>>(d, clk, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tf = (double)blocks * (threads / 32) * iters * 8 * 512 / (ms * 1e-3) / 1e12;
        printf("rep %2d: %.1f ms  %.2f TF/s  effective SM clock %.0f MHz\n", rep, ms, tf, (double)clk[1] / clk[0] * 1e3);
    }
    return 0;
}
