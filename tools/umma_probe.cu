// umma_probe.cu -- pins the tcgen05 conventions the split-precision kernel relies on, on real hardware:
// shared-memory descriptor (K-major, no swizzle: LBO = k-chunk distance, SBO = 8-row-group distance), instruction
// descriptor (BF16 x BF16 -> FP32, M = 128, N = 256), TMEM accumulator addressing and the 32x32b tcgen05.ld pattern,
// a bulk-copied operand next to an operand written with ordinary stores (+ fence.proxy.async), tcgen05.commit.
// One CTA computes D (128 x 256) = A (128 x KB) * B^T (256 x KB) with both operands split into two BF16 slices
// (three MMAs per k-step) and the host compares with a double-precision product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I channel-estimation_b200/csrc -o tools/bin/umma_probe tools/umma_probe.cu
//   tools/bin/umma_probe [swap]        (swap: exchange LBO and SBO -- must FAIL if the convention above is right)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "umma.cuh"

using namespace umma;

constexpr int M = 128, N = 256, KB = 64, NCH = KB / 8;       // KB k-elements = NCH chunks of 8
constexpr int A_SLICE = NCH * M * 16, B_SLICE = NCH * N * 16; // bytes of one slice of a tile

struct Shared {
    uint64_t full_a, full_b, done;
    uint32_t tmem_base;
    int abort_flag;
};

__global__ void __launch_bounds__(320, 1) k_probe(const uint8_t* __restrict__ a_image, const float* __restrict__ b_in,
                                                   float* __restrict__ d_out, int swap, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_s = smem;                       // [slice][chunk][row 128][16 B]
    uint8_t* b_s = smem + 2 * A_SLICE;         // [slice][chunk][row 256][16 B]
    __shared__ Shared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&sh.full_a, 1); mbar_init(&sh.full_b, 128); mbar_init(&sh.done, 1);
        sh.abort_flag = 0;
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&sh.tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;
    volatile int* ab = &sh.abort_flag;

    if (warp == 0) {
        if (lane == 0) {                                       // A: one bulk copy of the prepacked image
            mbar_arrive_expect_tx(&sh.full_a, 2 * A_SLICE);
            bulk_g2s(a_s, a_image, 2 * A_SLICE, &sh.full_a);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            bool ok = mbar_wait(&sh.full_a, 0, ab) && mbar_wait(&sh.full_b, 0, ab);
            if (ok) {
                tc_fence_after();
                const uint32_t idesc = idesc_bf16_f32(M, N);
                const uint32_t a0 = smem_u32(a_s), b0 = smem_u32(b_s);
                const uint32_t lbo_a = swap ? 128 : M * 16, sbo_a = swap ? M * 16 : 128;
                const uint32_t lbo_b = swap ? 128 : N * 16, sbo_b = swap ? N * 16 : 128;
                bool acc = false;
                for (int ks = 0; ks < KB / 16; ++ks)
                    for (int pr = 0; pr < 3; ++pr) {           // (Ah,Bh), (Ah,Bl), (Al,Bh)
                        const int sa = pr == 2 ? 1 : 0, sb = pr == 1 ? 1 : 0;
                        const uint64_t da = smem_desc(a0 + sa * A_SLICE + ks * 2 * M * 16, lbo_a, sbo_a);
                        const uint64_t db = smem_desc(b0 + sb * B_SLICE + ks * 2 * N * 16, lbo_b, sbo_b);
                        mma_bf16(tmem, da, db, idesc, acc);
                        acc = true;
                    }
            }
            mma_commit(&sh.done);                              // arrives once the MMAs above (if any) are complete
        }
    } else if (warp < 6) {                                     // B generators: thread n owns rows n and n + 128
        const int t = tid - 64;
        for (int half = 0; half < 2; ++half) {
            const int n = t + 128 * half;
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) split_bf16x2(b_in[n * KB + ch * 8 + 2 * q], b_in[n * KB + ch * 8 + 2 * q + 1], hi[q], lo[q]);
                *reinterpret_cast<uint4*>(b_s + (ch * N + n) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(b_s + B_SLICE + (ch * N + n) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        fence_proxy_async();
        mbar_arrive(&sh.full_b);
    } else {                                                   // epilogue: warp % 4 selects the lane quarter
        const int q = warp & 3;
        if (mbar_wait(&sh.done, 0, ab)) {
            tc_fence_after();
            for (int c0 = 0; c0 < N; c0 += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
                const int row = q * 32 + lane;
#pragma unroll
                for (int i = 0; i < 32; ++i) d_out[row * N + c0 + i] = v[i];
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (tid == 0) *status = sh.abort_flag;
    if (warp == 1) tmem_dealloc(tmem, 256);
}

static uint16_t f2bf(float x) {                                // round to nearest even
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float x; memcpy(&x, &u, 4); return x; }

int main(int argc, char** argv) {
    const int swap = argc > 1 && !strcmp(argv[1], "swap");
    std::vector<float> A(M * KB), B(N * KB);
    srand(7);
    for (auto& x : A) x = (float)rand() / RAND_MAX - 0.5f;
    for (auto& x : B) x = ((float)rand() / RAND_MAX - 0.5f) * (1.0f + 100.0f * ((float)rand() / RAND_MAX));
    // A image: [slice][chunk][row][8 bf16], built on the host exactly as the device setup packs W
    std::vector<uint16_t> img(2 * NCH * M * 8);
    for (int r = 0; r < M; ++r)
        for (int k = 0; k < KB; ++k) {
            const float x = A[r * KB + k];
            const uint16_t h = f2bf(x), l = f2bf(x - bf2f(h));
            img[((size_t)(0 * NCH + k / 8) * M + r) * 8 + k % 8] = h;
            img[((size_t)(1 * NCH + k / 8) * M + r) * 8 + k % 8] = l;
        }
    uint8_t* d_img; float *d_b, *d_d; int* d_status;
    cudaMalloc(&d_img, img.size() * 2); cudaMalloc(&d_b, B.size() * 4); cudaMalloc(&d_d, M * N * 4); cudaMalloc(&d_status, 4);
    cudaMemcpy(d_img, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(d_d, 0, M * N * 4);
    const int smem = 2 * A_SLICE + 2 * B_SLICE;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe<<<1, 320, smem>>>(d_img, d_b, d_d, swap, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("umma_probe: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> D(M * N); int status = 0;
    cudaMemcpy(D.data(), d_d, M * N * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost);
    double max_rel = 0, max_ref = 0;
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
            double ref = 0, mag = 0;
            for (int k = 0; k < KB; ++k) { ref += (double)A[r * KB + k] * B[n * KB + k]; mag += fabs((double)A[r * KB + k] * B[n * KB + k]); }
            max_rel = fmax(max_rel, fabs(D[r * N + n] - ref) / mag);
            max_ref = fmax(max_ref, fabs(ref));
        }
    printf("umma_probe%s: abort=%d  max |D - ref| / sum|a||b| = %.3e  (max |ref| %.3f)  -> %s\n", swap ? " [LBO/SBO swapped]" : "",
           status, max_rel, max_ref, (status == 0 && max_rel < 3e-5) ? "PASS" : "FAIL");
    return (status == 0 && max_rel < 3e-5) ? 0 : 1;
}
