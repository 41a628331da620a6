// dmma_probe.cu -- microbenchmark of the FP64 tensor pipe (DMMA m8n8k4) on sm_100a:
// throughput vs warps per SM and independent accumulator chains per warp, with and without
// interleaved DFMA / shared-memory fragment loads.  Development tool; results in profiles/.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH, int DF>
__global__ void k(double* out, int iters) {
    double c[CH][2], f[DF > 0 ? DF : 1];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < (DF > 0 ? DF : 1); ++i) f[i] = threadIdx.x + i;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < DF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < (DF > 0 ? DF : 1); ++i) s += f[i];
    if (s == 12345.678) out[0] = s;
}
template <int CH, int DF>
void run(int warps_per_sm, int nsm, double* d) {
    int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
    int blocks = nsm * (warps_per_sm / (threads / 32));
    int iters = 200000 / CH;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH, DF><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e0);
    k<CH, DF><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dm = (double)blocks * (threads / 32) * iters * CH;
    double tf = dm * 512 / (ms * 1e-3) / 1e12;
    double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * CH) * 1.0;   // cycles per DMMA per warp
    printf("warps/SM %2d chains %2d dfma %2d : %6.2f TF/s (DMMA only)  %.1f cyc/DMMA/warp  %.1f cyc/DMMA/SMSP\n",
           warps_per_sm, CH, DF, tf, cyc, cyc / (warps_per_sm / 4.0 > 1 ? warps_per_sm / 4.0 : 1));
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int nsm = pr.multiProcessorCount;
    double* d; cudaMalloc(&d, 64);
    int ws[] = {4, 8, 16, 32};
    for (int w : ws) { run<1, 0>(w, nsm, d); run<2, 0>(w, nsm, d); run<4, 0>(w, nsm, d); run<8, 0>(w, nsm, d); run<16, 0>(w, nsm, d); }
    for (int w : ws) { run<8, 4>(w, nsm, d); run<8, 8>(w, nsm, d); run<8, 16>(w, nsm, d); }
    return 0;
}
