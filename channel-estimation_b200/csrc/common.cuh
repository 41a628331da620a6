// common.cuh -- small device helpers shared by the kernels (complex FP64 arithmetic, FP64
// tensor-core MMA, counter-based RNG).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;   // (re, im)

__host__ __device__ __forceinline__ cplx cmake(double r, double i) { cplx c; c.x = r; c.y = i; return c; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { /* conj(a)*b */ return cmake(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// Complex division, Smith's method (the algorithm NumPy and most BLAS-level runtimes use).
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    if (fabs(b.x) >= fabs(b.y)) {
        double rat = b.y / b.x, scl = 1.0 / (b.x + b.y * rat);
        return cmake((a.x + a.y * rat) * scl, (a.y - a.x * rat) * scl);
    }
    double rat = b.x / b.y, scl = 1.0 / (b.y + b.x * rat);
    return cmake((a.x * rat + a.y) * scl, (a.y * rat - a.x) * scl);
}
// Reciprocal of a positive, normal double: single-precision seed (MUFU.RCP, 23 bits) and two Newton steps in FP64
// (x <- x (2 - d x): 46, then full precision; the last step leaves <= 1 ulp).  A handful of instructions instead of the
// ~30 of a correctly rounded FP64 division; used where the hot path divides once per equalised value.
__device__ __forceinline__ double rcp_pos(double d) {
    double x = (double)__frcp_rn((float)d);
    x = x * fma(-d, x, 2.0);
    x = fma(x, fma(-d, x, 1.0), x);
    return x;
}
// a / b for complex a, b as a conj(b) / |b|^2 with the fast reciprocal (one per value; |b|^2 stays far inside the float
// range for channel coefficients).  Agrees with Smith's division to a few ulp.
__device__ __forceinline__ cplx cdiv_fast(cplx a, cplx b) {
    const double inv = rcp_pos(fma(b.x, b.x, b.y * b.y));
    return cmake(fma(a.x, b.x, a.y * b.y) * inv, fma(a.y, b.x, -(a.x * b.y)) * inv);
}
// 16-byte load through the read-only path (data no thread of the kernel writes); not volatile, so the compiler may batch
// several of them ahead of the arithmetic and stores that follow
__device__ __forceinline__ cplx ld_nc(const cplx* ptr) {
    cplx r;
    asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(ptr));
    return r;
}
__device__ __forceinline__ double dneg(double x) {   // sign flip on the integer pipe
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}

// FP64 tensor-core MMA (DMMA): C(8x8) += A(8x4, row) * B(4x8, col).
// Fragment ownership, lane = 4*g + t (g = lane>>2, t = lane&3):
//   a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// CHEST_3M: complex tile products on the tensor pipe use three real multiplications instead of four
// ((ar+ai) br, -ai (br+bi), ar (bi-br); re = first + second, im = first + third).  The DMMA pipe is the
// bound of K2 / K4, so this removes a quarter of its work; the rounding error stays O(eps * sum|a||b|).
#ifndef CHEST_3M
#define CHEST_3M 1
#endif
// D = A * B + C with separate accumulator input (starts a new chain from another tile's partial sums)
__device__ __forceinline__ void dmma884c(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// complex tile update: (cr + j ci) += (ar + j ai) * (br + j bi), nbi = -bi
__device__ __forceinline__ void zmma884(double (&cr)[2], double (&ci)[2], double ar, double ai,
                                        double br, double bi, double nbi) {
    dmma884(cr[0], cr[1], ar, br);
    dmma884(cr[0], cr[1], ai, nbi);
    dmma884(ci[0], ci[1], ar, bi);
    dmma884(ci[0], ci[1], ai, br);
}

// ---------------------------------------------------------------- Philox4x32-10
struct Philox4 { uint32_t v[4]; };
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}
// 53-bit uniform in (0,1) from two 32-bit words
__host__ __device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
    uint64_t w = ((uint64_t)hi << 32) | lo;
    return (double)(w >> 11) * (1.0 / 9007199254740992.0) + (0.5 / 9007199254740992.0);
}
// RNG streams (counter word c2 = stream | snr << 8)
enum { RS_DOPPLER = 0, RS_PHASE = 1, RS_BITS0 = 2 /* +scheme */, RS_PILOT0 = 5 /* +waveform */, RS_NOISE = 7, RS_CHAN_GAUSS = 8, RS_SV_H = 9 };
