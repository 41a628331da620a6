// chest_api.cu -- context, one-time setup and the extern "C" launchers declared in
// include/chest_b200.h.  Host code only orchestrates: every arithmetic step of the loop body
// runs in the kernels of kernels.cuh on the context's stream.  No CPU fallback.
#include "../../include/chest_b200.h"
#include "kernels.cuh"
#include "kernels_tc.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(CHEST_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)
#define ARG(cond)                                                                                  \
    do {                                                                                           \
        if (!(cond)) return fail(CHEST_ERR_ARG, std::string("argument check failed: ") + #cond);   \
    } while (0)

namespace {

// Owning device buffer: move-only, freed by its destructor (so every DevBuf member of Ctx / Waveform / Scheme is
// released by `delete ctx` -- nothing to list by hand, nothing to forget).
template <class T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const T* h, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    cudaError_t upload(const std::vector<T>& v, cudaStream_t s) { return upload(v.data(), v.size(), s); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct Waveform {
    bool set = false;
    int K = 0;
    DevBuf<cplx> G, Q, Gt;
    DevBuf<cplx> Q1; DevBuf<double> Q2;                // three-multiplication planes of Q^H: (re, re - im), im; row stride Np
    DevBuf<cplx> Gt1; DevBuf<double> Gt2;              // planes of G (not conjugated), rows = samples: (re, re + im), -im; row stride Kp
    bool det_on = false; DevBuf<uint8_t> zw_g; int zw_stride = 0;   // FBMC perfect-CSI columns equalised / detected by k_perfect_fbmc_det
    bool twin_on = false;                              // the perfect-CSI twin of this waveform runs in k_perfect_twin_fbmc (no PERF units)
    int pf_state = 0;                                  // polyphase perfect-CSI pass: 0 unchecked, 1 usable, -1 not (no modem description, mismatch, too large)
    DevBuf<int2> pf_groups; int pf_n_groups = 0;
    DevBuf<cplx> f_r1; DevBuf<double> f_r2;            // planes of r = H s of the factored perfect-CSI pass: (re, re + im), im - re; [column][Np]
    DevBuf<cplx> HG1; DevBuf<double> HG2;              // H*G planes: (re, re + im), im - re; [rep][K][Np]
    DevBuf<int> q_klo, q_khi, gt_klo, gt_khi, hg_klo, hg_khi, d_jlo, d_jhi;
    DevBuf<int> q8_klo, q8_khi, hg8_klo, hg8_khi, gt8_klo, gt8_khi;
    DevBuf<int2> pairs; int n_pairs = 0;               // (row tile, column tile) pairs of D with overlapping supports      // per 8 columns of Q / of H*G (warp-level clipping in K2 / K3a)
    double d_struct_pairs = 0;   // (i,j) pairs of D inside the structural support
    std::vector<int> g_lo, g_hi, q_lo, q_hi;
    std::vector<double> hg_rows;   // per column: rows of H*G written by k_apply_hg
    int nsch = 0; int sch[2] = {0, 0};
    // batch state
    DevBuf<cplx> x, s, r0, y, D, htrue;
    // factored perfect-CSI pass (perf_mode 1): column tables of the units' v / y_ic slots, and the two vector sets
    DevBuf<int> q_lo_d, q_hi_d;
    DevBuf<cplx> Pd; DevBuf<int> pd_klo, pd_khi; bool pd_built = false;   // diag(D) operand: P[i][t*N + n] = conj(Q[n,i]) G[n - tau_t, i]
    std::vector<cplx> Gh, Qh;                           // host copies of G, Q (finalize builds Pd from them)
    DevBuf<int64_t> f_voff, f_yoff; DevBuf<int> f_rep; DevBuf<cplx> f_s, f_r;
    int f_cols = 0, perf_base = 0, perf_nblk = 0;
    // FFT modem (chest_set_modem): plan, tables and the per-symbol IFFT outputs of the FBMC modulator
    bool modem_set = false; ModemDev modem{};
    DevBuf<int> m_bin; DevBuf<double> m_filt; DevBuf<cplx> m_phase, m_tw, m_Z, m_in, m_out;
    int d_alloc_batch = 0;      // batch size D / HG1 / HG2 are allocated for (0: not yet)
    // device-side setup (chest_setup_correlations): thresholded R_Dij_hP of all pilots, row-tile-major like D
    DevBuf<cplx> Rsup; int rsup_P = 0; double rsup_thr = 0;
    // factored estimated-CSI cancellation (k_est_channel / k_est_factored): the pseudo-channels M_q of the pilots [P][T][N] kept from
    // chest_setup_correlations, whether any wrapped (corner) entry is non-zero, the largest |R_Dij_hP| entry the threshold removed,
    // and the EST units of this waveform's factored schemes
    DevBuf<cplx> Mq; int mq_P = 0; bool mq_corner = false; double rsup_zeroed_max = 0;
    DevBuf<int4> ef_desc; int ef_n_units = 0;
    DevBuf<int> g_lo_d, g_hi_d;
    int tile = 64;              // CTA tile size of the GEMMs on this waveform (48 or 64)
    double flops_d = 0, flops_demod = 0, flops_mod = 0;
};
struct Constellation {
    bool set = false;
    int order = 0, nbits = 0, n_axis = 0, is_qam = 0;
    bool real = false;           // every symbol (hence every unit-modulus pilot) is exactly real
    DevBuf<cplx> symbol, pilot; DevBuf<double> level; DevBuf<int> word_of_grid;
    ConstDev dev{};
};
struct MmseVariant {
    bool set = false;
    int n_tiles = 0;
    DevBuf<int> tile_ptr, tile_delta;
    std::vector<DevBuf<cplx>> frag, diag;
    DevBuf<cplx> diag_frag;        // [snr][rt][pq][32 lanes]
    DevBuf<WTiles> table;
    int64_t nnz_offdiag_pairs = 0;
    DevBuf<cplx> rinv; bool rinv_set = false; double w_zeroed_max = 0;   // pinv(R_hP_est) [snr][P x P] of chest_build_mmse; largest |W| entry its threshold removed
    // split-BF16 tensor-core mode (chest_set_precision): active columns per 128-row tile and the operand images
    DevBuf<int> tc_jlist, tc_jptr; DevBuf<uint8_t> tc_img; int tc_entries = 0, tc_row_tiles = 0; bool tc_packed = false;
    std::vector<int> tc_jptr_h;
};
struct Scheme {
    bool set = false;
    int waveform = 0, K = 0, K_in = 0, P = 0, n_data = 0, detect = 0, constellation = 0, n_bits = 0;
    double kappa = 1, dpr = 1;
    DevBuf<int> c_rowptr, c_col, ct_colptr, ct_row, pilot_pos, data_pos, pos2data, row_col0, long_rows, lr_ptr, lr_kcol;
    DevBuf<cplx> lr_frag; int n_lr_tiles = 0;
    DevBuf<cplx> row_val0; int n_long_rows = 0;
    DevBuf<cplx> c_val, ct_val;
    DevBuf<uint32_t> edge_mask;
    std::vector<uint8_t> considered;
    MmseVariant mm[2];
    DevBuf<cplx> xP, hP, hdiag, xD[2];
    DevBuf<uint32_t> txword, txw_t;
    DevBuf<uint8_t> bits;
    DevBuf<int4> rowinfo; DevBuf<int> multi_d; int n_multi = 0; bool fuse_ok = false;
    DevBuf<cplx> interp;           // SV.m chain: interpolation matrix K x P, row-major (chest_set_interpolation)
    int64_t n_bits_edge = 0;
    int64_t c_nnz = 0;
    bool c_real = false;
};

struct Ctx {
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    // channel
    bool chan_set = false;
    int N = 0, Lt = 0, T = 0, paths = 0, model = 0;
    double fD = 0, dt = 0;
    std::vector<int> tap_delay; std::vector<double> tap_amp;
    DevBuf<int> d_tap_delay; DevBuf<double> d_tap_amp;
    // discrete Doppler spectrum (FF.m:151-178): shifts -n_shift..n_shift, coefficient sqrt(spectrum * pdp / 2) per (bin, tap)
    int n_shift = 0; std::vector<double> dspec; DevBuf<double> d_dcoef; DevBuf<cplx> chan_gauss;
    Waveform wf[2];
    Constellation cst[2];
    Scheme sch[3];
    int S = 0; std::vector<double> pn; DevBuf<double> d_noise_scale;
    // batch
    bool finalized = false;
    int max_batch = 0, cur_batch = 0, last_iter = 0;
    DevBuf<double> doppler_u, phase_u; DevBuf<cplx> noise, h;
    DevBuf<int32_t> pilot_idx[2];
    // chest_prefetch_draws: two library-owned copies of the draw buffers filled on a copy stream
    struct Prefetch {
        DevBuf<double> du, pu; DevBuf<cplx> noise, cg; DevBuf<uint8_t> bits[3]; DevBuf<int32_t> pidx[2];
        cudaEvent_t landed = nullptr, released = nullptr; bool used = false;
    } pf[2];
    int pf_next = 0;
    int precision = 0;           // 0: FP64 DMMA everywhere (default); 1: estimated-CSI cancellation on tcgen05, split BF16 (kernels_tc.cuh)
    int tc_p8 = 0;               // pilots per scheme padded to the k-chunking of the tensor-core mode (16 or 32)
    DevBuf<TcItem> tc_items; DevBuf<int> tc_cta_ptr, tc_status; int tc_n_items = 0, tc_grid = 0;
    double tc_mma_flops = 0;     // dense BF16 flops one launch of k_ic_est_tc executes (roofline of the reduced-precision mode)
    int perf_mode = 1;           // 1 (default): factored, y - Q^H H (G v) + h v, D never formed; 0: D materialised (K2) and applied densely
    int n_est_units = 0;
    // estimated-CSI cancellation: 0 auto (factored form only for schemes where it equals the tile form of W to rounding), 1 tiles
    // (always the thresholded W), 2 factored (every scheme that has the factors: the stated-tolerance mode, chest_set_estimator_mode)
    int est_mode = 0; bool est_fact[3] = {false, false, false};
    bool chain_cm = true, chain_cm_perf = true;   // unit scratch of the chain kernels' columns is column-major (IcParams::chain_colmajor): factored EST units / detected PERF units
    DevBuf<cplx> hest; DevBuf<unsigned long long> zmax;
    cudaEvent_t ev_ef[18] = {}; float est_fact_ms = 0;
    cudaStream_t copy_stream = nullptr;
    DevBuf<uint32_t> err;
    DevBuf<cplx> scratch, tmp_a, tmp_b;
    DevBuf<IcCta> ctas; int n_ctas = 0, ctas_for_batch = -1;
    int K_max = 0;
    int64_t launches = 0;
    bool profiling = false;
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_hg[4] = {};   // k_apply_hg of the two waveforms (profiling)
    cudaEvent_t ev_gd[2] = {};   // end of k_gemm_d of the two waveforms
    cudaEvent_t ev_ic[36] = {};  // after every k_ic_main (+ factored chain) / k_ic_light launch
    cudaEvent_t ev_mn[18] = {};  // after k_ic_main alone (before the factored perfect-CSI chain)
    cudaEvent_t ev_tw[2] = {};   // around the fused perfect-CSI twin kernels
    cudaEvent_t ev_k1[8] = {};   // inside stage 1: after k_synth_h, k_tx_symbols, per waveform after s = G x and after r0 = H s
    float kernel_ms[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // k_apply_hg, k_gemm_d, k_ic_main, k_ic_light, factored perfect-CSI chain, diag(D) GEMM, k_synth_h, k_tx_symbols, s = G x, k_apply_h
    float hg_ms = 0; double hg_bytes = 0;
    cudaEvent_t user_ev[4] = {};
    float stage_ms[7] = {0, 0, 0, 0, 0, 0, 0};
    DevBuf<double> probe;
    DevBuf<unsigned int> queue; int ic_grid = 0, ic_light_grid = 0, ic_cfg = -1; size_t ic_smem = 0;
    DevBuf<unsigned long long> trace;
    // asynchronous runs (chest_run_batch_async / chest_wait): pinned host copy of the counters, pending-run state
    uint32_t* err_pinned = nullptr; size_t err_pinned_n = 0;
    bool pending = false; size_t pending_n_err = 0; bool trace_on = false; int pending_iter = 0, pending_rep = 0;
    DevBuf<unsigned long long> totals;    // [snr][it][12] sums over realizations (chest_multi_run)
    DevBuf<cplx> sv_h, sv_noise; DevBuf<double> sv_pn; DevBuf<uint32_t> sv_err;   // chest_sv_run_batch
    bool mse_on = false; DevBuf<double> mse;                      // chest_set_mse_accumulation: [rep][snr][it][scheme]
    DevBuf<int> setup_pil; DevBuf<double> setup_rt, setup_tp; DevBuf<cplx> setup_corner, setup_rhp, setup_eye;   // chest_setup_correlations scratch
    DevBuf<cplx> setup_rinv; DevBuf<int> setup_mask, setup_trt;   // chest_build_mmse scratch, kept across calls (a velocity sweep rebuilds W per velocity)
};

Ctx* from(uint64_t h) { return reinterpret_cast<Ctx*>(static_cast<uintptr_t>(h)); }

void support_ranges(const double* A, int N, int K, std::vector<int>& lo, std::vector<int>& hi) {
    lo.assign(K, N); hi.assign(K, 0);
    for (int j = 0; j < K; ++j) {
        const double* c = A + 2 * (size_t)N * j;
        int a = 0, b = N;
        while (a < N && c[2 * a] == 0.0 && c[2 * a + 1] == 0.0) ++a;
        while (b > a && c[2 * (b - 1)] == 0.0 && c[2 * (b - 1) + 1] == 0.0) --b;
        lo[j] = a; hi[j] = b;
        if (a >= b) { lo[j] = 0; hi[j] = 0; }
    }
}
void tile_ranges(const std::vector<int>& lo, const std::vector<int>& hi, int tile, int extra, int cap,
                 std::vector<int>& tlo, std::vector<int>& thi) {
    int n = (int)lo.size(), nt = (n + tile - 1) / tile;
    tlo.assign(nt, cap); thi.assign(nt, 0);
    for (int j = 0; j < n; ++j) {
        if (hi[j] <= lo[j]) continue;
        int t = j / tile;
        tlo[t] = std::min(tlo[t], lo[j]);
        thi[t] = std::max(thi[t], std::min(cap, hi[j] + extra));
    }
    for (int t = 0; t < nt; ++t) if (thi[t] <= tlo[t]) { tlo[t] = 0; thi[t] = 0; }
}

template <int MODE, int WM, int WN, int TMW>
cudaError_t launch_gemm_geo(Ctx* c, const GemmParams& p, int n_z) {
    constexpr int TM = 8 * TMW * WM, TN = 16 * WN;
    constexpr int smem = 3 * (TM + TN) * 20 * (int)sizeof(cplx);     // 2 stages x (A + B) rows x (16+4) x (complex + double plane)
    static bool attr_dev[64] = {};                                   // function attributes are per device
    bool& attr_done = attr_dev[c->device & 63];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm<MODE, WM, WN, TMW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((p.M + TM - 1) / TM, (p.n_cols + TN - 1) / TN, n_z);
    k_gemm<MODE, WM, WN, TMW><<<grid, 32 * WM * WN, smem, c->stream>>>(p);
    c->launches++;
    return cudaGetLastError();
}
template <int WM, int WN, int TMW>
cudaError_t launch_gemm_d_geo(Ctx* c, const GemmDParams& p) {
    constexpr int TM = 8 * TMW * WM, TN = 16 * WN;
    constexpr int smem = GEMMD_STAGES * 3 * (TM + TN) * (GEMMD_KT + 4) * (int)sizeof(double);   // stages x rows x (KT+4) x (complex + double plane)
    static int grid_dev[64] = {};                                    // function attributes and occupancy are per device
    int& grid = grid_dev[c->device & 63];
    if (!grid) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_d<WM, WN, TMW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gemm_d<WM, WN, TMW>, 32 * WM * WN, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        grid = per_sm * c->n_sm;                                    // persistent: one resident wave
    }
    const long long total = (long long)p.n_pairs * p.n_rep;
    if (total == 0) return cudaSuccess;
    if (total > 0x7fffffffLL) return cudaErrorInvalidValue;
    k_gemm_d<WM, WN, TMW><<<(int)std::min<long long>(grid, total), 32 * WM * WN, smem, c->stream>>>(p);
    c->launches++;
    return cudaGetLastError();
}
// tile = 64 or 48 (square CTA tiles); the per-tile k-range tables must have been built for the same size
template <int MODE>
cudaError_t launch_gemm(Ctx* c, const GemmParams& p, int n_z, int tile) {
    if (tile == 48) return launch_gemm_geo<MODE, 2, 3, 3>(c, p, n_z);
    return launch_gemm_geo<MODE, 2, 4, 4>(c, p, n_z);
}

SchemeDev scheme_dev(Ctx* c, int si) {
    Scheme& s = c->sch[si];
    Waveform& w = c->wf[s.waveform];
    SchemeDev d{};
    d.waveform = s.waveform; d.K = s.K; d.K_in = s.K_in; d.P = s.P; d.P4 = (s.P + 3) / 4;
    d.n_data = s.n_data; d.nbits = c->cst[s.constellation].nbits; d.detect_mode = s.detect;
    d.constellation = s.constellation; d.n_bits_total = s.n_bits;
    d.sqrt_kappa = std::sqrt(s.kappa); d.dpr = s.dpr; d.sqrt_dpr = std::sqrt(s.dpr);
    d.inv_sqrt_dpr = 1.0 / d.sqrt_dpr; d.inv_dpr = 1.0 / s.dpr;
#ifndef CHEST_VREAL
#define CHEST_VREAL 1
#endif
    d.v_real = (CHEST_VREAL && s.c_real && c->cst[s.constellation].real) ? 1 : 0;
    d.c_rowptr = s.c_rowptr.p; d.c_col = s.c_col.p; d.c_val = s.c_val.p;
    d.ct_colptr = s.ct_colptr.p; d.ct_row = s.ct_row.p; d.ct_val = s.ct_val.p;
    d.pilot_pos = s.pilot_pos.p; d.data_pos = s.data_pos.p; d.pos2data = s.pos2data.p; d.edge_mask = s.edge_mask.p;
    d.row_col0 = s.row_col0.p; d.row_val0 = s.row_val0.p; d.long_rows = s.long_rows.p; d.n_long_rows = s.n_long_rows;
    d.n_lr_tiles = s.n_lr_tiles; d.lr_ptr = s.lr_ptr.p; d.lr_kcol = s.lr_kcol.p; d.lr_frag = s.lr_frag.p;
    d.wdiag_frag[0] = s.mm[0].diag_frag.p; d.wdiag_frag[1] = s.mm[1].diag_frag.p;
    d.rowinfo = s.rowinfo.p; d.multi_d = s.multi_d.p; d.n_multi = s.n_multi; d.fuse_ok = s.fuse_ok ? 1 : 0;
    d.txw_t = s.txw_t.p; d.txw_t_w = s.txw_t.p;
    for (int v = 0; v < 2; ++v) {
        d.tile_ptr[v] = s.mm[v].tile_ptr.p; d.tile_delta[v] = s.mm[v].tile_delta.p; d.w[v] = s.mm[v].table.p;
    }
    int g = (w.sch[0] == si) ? 0 : 1;
    int B = c->cur_batch;
    d.xP = s.xP.p; d.txword = s.txword.p;
    d.x = w.x.p + (size_t)g * B * s.K;
    d.y = w.y.p + (size_t)g * c->S * B * s.K;
    d.hP = s.hP.p; d.hdiag = s.hdiag.p; d.xD[0] = s.xD[0].p; d.xD[1] = s.xD[1].p;
    return d;
}

int check_ready(Ctx* c) {
    if (!c) return fail(CHEST_ERR_ARG, "null handle");
    if (!c->finalized) return fail(CHEST_ERR_STATE, "context not finalized (call chest_finalize)");
    return CHEST_OK;
}

// ---------------------------------------------------------------- pipeline stages (all async on c->stream)
int stage_channel_discrete(Ctx* c, int n_rep, const cplx* gauss_dev) {
    dim3 grid((c->N + 127) / 128, c->T, n_rep);
    k_synth_h_discrete<<<grid, 128, 0, c->stream>>>(c->h.p, gauss_dev, c->d_dcoef.p, c->N, c->T, c->n_shift);
    c->launches++;
    CK(cudaGetLastError());
    return CHEST_OK;
}
int gen_chan_gauss(Ctx* c, int n_rep, uint64_t seed, int64_t first_rep) {
    const int n = (2 * c->n_shift + 1) * c->T;
    dim3 g((n + 63) / 64, n_rep);
    k_rng_cnormal<<<g, 64, 0, c->stream>>>(c->chan_gauss.p, n, n_rep, RS_CHAN_GAUSS, seed, first_rep);
    c->launches++;
    CK(cudaGetLastError());
    return CHEST_OK;
}

int stage_channel(Ctx* c, int n_rep, const double* du, const double* pu) {
    dim3 grid((c->N + SYNTH_THREADS * SYNTH_SEG - 1) / (SYNTH_THREADS * SYNTH_SEG), c->T, n_rep);
    k_synth_h<<<grid, SYNTH_THREADS, 4 * c->paths * sizeof(double), c->stream>>>(
        c->h.p, du, pu, c->d_tap_amp.p, c->N, c->T, c->paths, c->fD, c->dt, c->model);
    c->launches++;
    CK(cudaGetLastError());
    return CHEST_OK;
}

// D (row-tile-major) and the H*G operand planes for the whole batch; allocated by the first call that needs them
// (dense mode, chest_transmission_matrix) -- the factored loop body never does.
int ensure_d_buffers(Ctx* c, Waveform& w) {
    const int B = c->max_batch, N = c->N;
    if (w.d_alloc_batch == B) return CHEST_OK;
    const size_t n_d = (size_t)B * (((w.K + 7) / 8) * 8) * w.K, n_hg = (size_t)B * w.K * ((N + 1) & ~1);
    CK(w.D.alloc(n_d)); CK(w.HG1.alloc(n_hg)); CK(w.HG2.alloc(n_hg));
    // tiles of D without support overlap are never written by K2: they stay at this zero
    CK(cudaMemsetAsync(w.D.p, 0, n_d * sizeof(cplx), c->stream));
    CK(cudaMemsetAsync(w.HG1.p, 0, n_hg * sizeof(cplx), c->stream));          // pad element of odd N stays zero
    CK(cudaMemsetAsync(w.HG2.p, 0, n_hg * sizeof(double), c->stream));
    w.d_alloc_batch = B;
    return CHEST_OK;
}

// diag(D) as a GEMM over realizations: h[rep][i] = sum_{t,n} P[i][t*N + n] h[rep][t][n], P = conj(Q) shift_t(G).
// K x T x N complex: built on the first factored-mode run only (a waveform used for K1 / K2 alone never needs it).
int ensure_pd(Ctx* c, Waveform& w) {
    if (w.pd_built) return CHEST_OK;
    const int N = c->N, K = w.K, T = c->T;
    std::vector<cplx> P((size_t)K * T * N, cmake(0.0, 0.0));
    for (int i = 0; i < K; ++i)
        for (int t = 0; t < T; ++t) {
            const int d = c->tap_delay[t];
            for (int n = std::max(w.q_lo[i], d); n < w.q_hi[i]; ++n) {
                const cplx q = w.Qh[(size_t)n + (size_t)N * i], g = w.Gh[(size_t)(n - d) + (size_t)N * i];
                P[((size_t)i * T + t) * N + n] = cmake(q.x * g.x + q.y * g.y, q.x * g.y - q.y * g.x);
            }
        }
    CK(w.Pd.upload(P, c->stream));
    std::vector<int> ql, qh;
    tile_ranges(w.q_lo, w.q_hi, w.tile, 0, N, ql, qh);
    for (size_t a = 0; a < ql.size(); ++a) if (qh[a] > ql[a]) qh[a] += (T - 1) * N;   // the range spans all tap segments
    CK(w.pd_klo.upload(ql, c->stream)); CK(w.pd_khi.upload(qh, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    w.pd_built = true;
    return CHEST_OK;
}

int stage_transmission_matrix(Ctx* c, int wfi, int n_rep, int rep0) {
    Waveform& w = c->wf[wfi];
    { int rc_ = ensure_d_buffers(c, w); if (rc_) return rc_; }
    GemmDParams p{};
    const int Np = (c->N + 1) & ~1;                                 // even row stride of the operand planes
    p.M = w.K; p.n_cols = w.K; p.lda = Np; p.ldb = Np; p.n_rep = n_rep; p.rep0 = rep0;
    p.n_pairs = w.n_pairs; p.pairs = w.pairs.p;
    p.At1 = w.Q1.p; p.At2 = w.Q2.p; p.b1 = w.HG1.p; p.b2 = w.HG2.p;
    p.mt_klo = w.q_klo.p; p.mt_khi = w.q_khi.p; p.nt_klo = w.hg_klo.p; p.nt_khi = w.hg_khi.p;
    p.m8_klo = w.q8_klo.p; p.m8_khi = w.q8_khi.p; p.n8_klo = w.hg8_klo.p; p.n8_khi = w.hg8_khi.p;
    p.out = w.D.p; p.hdiag = w.htrue.p;
    dim3 ghg((w.K + HG_COLS - 1) / HG_COLS, n_rep);
    if (c->profiling && n_rep > 1) CK(cudaEventRecord(c->ev_hg[2 * wfi], c->stream));
    k_apply_hg<<<ghg, 128, 0, c->stream>>>(w.HG1.p, w.HG2.p, w.G.p, c->h.p, c->d_tap_delay.p, w.hg_klo.p, w.hg_khi.p,
                                           c->N, Np, w.K, c->T, rep0, w.tile);
    c->launches++;
    CK(cudaGetLastError());
    if (c->profiling && n_rep > 1) CK(cudaEventRecord(c->ev_hg[2 * wfi + 1], c->stream));
    if (w.tile == 48) CK((launch_gemm_d_geo<2, 3, 3>(c, p)));
    else CK((launch_gemm_d_geo<2, 4, 4>(c, p)));
    if (c->profiling && n_rep > 1) CK(cudaEventRecord(c->ev_gd[wfi], c->stream));
    return CHEST_OK;
}

// Perfect-CSI cancellation without D: y_ic = y - Q^H (H (G v)) + h v for every (realization, scheme, SNR) column
// (DS.m:541-543 with D = Q^H H G, h = diag D).  Two support-aware GEMMs over all columns of the batch with the banded
// channel between them; the second GEMM's epilogue writes y_ic into the units' scratch.
// device-resident core: x_dev [n_cols][L*K] -> s_dev [n_cols][N]
int modem_modulate_dev(Ctx* c, Waveform& w, const cplx* x_dev, int n_cols, cplx* s_dev) {
    const ModemDev& md = w.modem;
    const size_t smem = (size_t)3 * md.nfft * sizeof(cplx);
    const int thr = std::min(256, std::max(32, ((md.nfft / 2 + 31) / 32) * 32));
    dim3 g(md.Ksym, n_cols);
    if (md.kind == 0) {
        CK(w.m_Z.alloc((size_t)n_cols * md.Ksym * md.nfft));
        k_modem_ifft<<<g, thr, smem, c->stream>>>(md, x_dev, w.m_Z.p, nullptr);
        dim3 g2(n_cols, (md.N + 255) / 256);
        k_fbmc_overlap_add<<<g2, 256, 0, c->stream>>>(md, w.m_Z.p, s_dev);
        c->launches += 2;
    } else {
        CK(cudaMemsetAsync(s_dev, 0, sizeof(cplx) * (size_t)n_cols * md.N, c->stream));     // zero guard samples
        k_modem_ifft<<<g, thr, smem, c->stream>>>(md, x_dev, nullptr, s_dev);
        c->launches++;
    }
    CK(cudaGetLastError());
    return CHEST_OK;
}
int modem_demodulate_dev(Ctx* c, Waveform& w, const cplx* r_dev, int n_cols, cplx* y_dev) {
    const ModemDev& md = w.modem;
    const size_t smem = (size_t)3 * md.nfft * sizeof(cplx);
    const int thr = std::min(256, std::max(32, ((md.nfft / 2 + 31) / 32) * 32));
    dim3 g(md.Ksym, n_cols);
    k_modem_fft<<<g, thr, smem, c->stream>>>(md, r_dev, y_dev);
    c->launches++;
    CK(cudaGetLastError());
    return CHEST_OK;
}

#ifndef PERF_FBMC_CW
#define PERF_FBMC_CW 1
#endif
// Can the perfect-CSI pass of this waveform run through the polyphase modem (k_perfect_fbmc)?  Needs an FBMC description
// from chest_set_modem next to the dense matrices, shared-memory room for CW columns, and -- since the two descriptions come
// from the caller independently -- the modem must reproduce G and Q^H on probe vectors.
int check_polyphase_pass(Ctx* c, Waveform& w) {
    if (w.pf_state) return CHEST_OK;
    w.pf_state = -1;
    if (getenv("CHEST_CHAIN_GEMM")) return CHEST_OK;            // development: force the GEMM chain
    const ModemDev& md = w.modem;
    const int N = c->N, K = w.K;
    if (!w.modem_set || !w.set || md.L * md.Ksym != K || md.N != N) return CHEST_OK;
    if (md.kind == 0 && md.time_spacing * 2 != md.nfft) return CHEST_OK;
    const size_t nx = std::max((size_t)md.Ksym * md.nfft, md.kind == 0 ? (size_t)0 : (size_t)N);
    if (md.kind == 0 && (size_t)N > nx) return CHEST_OK;
    for (int q = 0; q < md.plan.n_stage; ++q) if (md.plan.radix[q] > 7) return CHEST_OK;     // the batched FFT carries radices 2, 3, 4, 5, 7
    const size_t smem = ((size_t)2 * PERF_FBMC_CW * nx + md.nfft) * sizeof(cplx) + (size_t)md.Np * sizeof(double) + (size_t)md.L * sizeof(int);
    if (smem > 200 * 1024) return CHEST_OK;
    // probes: three unit symbol vectors through the modulator against the columns of G, three unit samples through the
    // demodulator against the rows of Q^H
    const int cols[3] = {0, K / 2 + 1 < K ? K / 2 + 1 : 0, K - 1}, rows[3] = {w.q_lo[K / 2], (w.q_lo[K / 2] + w.q_hi[K / 2]) / 2, std::max(0, w.q_hi[K - 1] - 1)};
    std::vector<cplx> x((size_t)3 * K, cmake(0.0, 0.0)), e((size_t)3 * N, cmake(0.0, 0.0)), s_h((size_t)3 * N), y_h((size_t)3 * K);
    for (int q = 0; q < 3; ++q) { x[(size_t)q * K + cols[q]] = cmake(1.0, 0.0); e[(size_t)q * N + rows[q]] = cmake(1.0, 0.0); }
    DevBuf<cplx> dx, de, ds, dy;
    CK(dx.upload(x, c->stream)); CK(de.upload(e, c->stream)); CK(ds.alloc((size_t)3 * N)); CK(dy.alloc((size_t)3 * K));
    int rc = modem_modulate_dev(c, w, dx.p, 3, ds.p); if (rc) return rc;
    rc = modem_demodulate_dev(c, w, de.p, 3, dy.p); if (rc) return rc;
    CK(cudaMemcpyAsync(s_h.data(), ds.p, sizeof(cplx) * s_h.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(y_h.data(), dy.p, sizeof(cplx) * y_h.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    double dev = 0, mag = 0;
    for (int q = 0; q < 3; ++q) {
        for (int n = 0; n < N; ++n) {
            const cplx g = w.Gh[(size_t)n + (size_t)N * cols[q]], m = s_h[(size_t)q * N + n];
            dev = std::max(dev, std::hypot(g.x - m.x, g.y - m.y)); mag = std::max(mag, std::hypot(g.x, g.y));
        }
        for (int i = 0; i < K; ++i) {                           // (Q^H e_n)[i] = conj(Q[n, i])
            const cplx qv = w.Qh[(size_t)rows[q] + (size_t)N * i], m = y_h[(size_t)q * K + i];
            dev = std::max(dev, std::hypot(qv.x - m.x, -qv.y - m.y)); mag = std::max(mag, std::hypot(qv.x, qv.y));
        }
    }
    if (dev <= 1e-12 * mag) {
        w.pf_state = md.kind == 0 ? 1 : 2;                      // 2: CP-OFDM -- the one-column kernels (k_perfect_fbmc_det, k_demod_fbmc) only
        static size_t attr_dev[64] = {};                        // the attribute is per kernel and device: keep the largest need seen
        size_t& attr_smem = attr_dev[c->device & 63];
        if (smem > attr_smem) {
            attr_smem = smem;
            const int smem_e = (int)(smem + CHAIN24_NE * PERF_FBMC_THREADS * sizeof(cplx) + 16);      // + the e = y + h v buffer of the 24-point chain
            CK(cudaFuncSetAttribute(k_perfect_fbmc<PERF_FBMC_CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(k_demod_fbmc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(k_perfect_fbmc_det<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(k_perfect_fbmc_det<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_e));
            CK(cudaFuncSetAttribute(k_est_factored<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(k_est_factored<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_e));
        }
    }
    return CHEST_OK;
}

template <int WM, int WN, int TMW, bool BG, int EPI>
cudaError_t launch_gemm_ring_geo(Ctx* c, GemmRingParams& p) {
    constexpr int TM = 8 * TMW * WM, TN = 16 * WN;
    constexpr int smem = GEMMD_STAGES * 3 * (TM + TN) * (GEMMD_KT + 4) * (int)sizeof(double);
    static int grid_dev[64] = {};                                    // function attributes and occupancy are per device
    int& grid = grid_dev[c->device & 63];
    if (!grid) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_ring<WM, WN, TMW, BG, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gemm_ring<WM, WN, TMW, BG, EPI>, 32 * WM * WN, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        grid = per_sm * c->n_sm;
    }
    p.n_mt = (p.M + TM - 1) / TM; p.n_nt = (p.n_cols + TN - 1) / TN;
    const long long total = (long long)p.n_mt * p.n_nt;
    if (total == 0) return cudaSuccess;
    if (total > 0x7fffffffLL) return cudaErrorInvalidValue;
    k_gemm_ring<WM, WN, TMW, BG, EPI><<<(int)std::min<long long>(grid, total), 32 * WM * WN, smem, c->stream>>>(p);
    c->launches++;
    return cudaGetLastError();
}
template <bool BG, int EPI>
cudaError_t launch_gemm_ring(Ctx* c, GemmRingParams& p, int tile) {
    if (tile == 48) return launch_gemm_ring_geo<2, 3, 3, BG, EPI>(c, p);
    return launch_gemm_ring_geo<2, 4, 4, BG, EPI>(c, p);
}

// Can the estimated-CSI cancellation of scheme si run in factored form (kernels.cuh, k_est_factored)?  Needs the factored perfect-CSI
// mode (its scratch layout), the factors kept by the device-side setup (pseudo-channels of the pilots, pinv(R_hP_est) of both
// variants), no wrapped pseudo-channel entries, a modem description that reproduces G / Q and fits one column per CTA -- and either
// the explicit stated-tolerance mode or the proof that the thresholds of DS.m:263-264, 287-289 removed nothing but rounding noise.
int est_factored_usable(Ctx* c, int si, bool& on) {
    on = false;
    Scheme& s = c->sch[si];
    if (!s.set || c->perf_mode != 1 || c->est_mode == 1 || getenv("CHEST_LIGHT")) return CHEST_OK;
    Waveform& w = c->wf[s.waveform];
    if (!w.Mq.p || w.mq_P != s.P || w.mq_corner || !s.mm[0].rinv_set || !s.mm[1].rinv_set || !s.mm[0].set || !s.mm[1].set) return CHEST_OK;
    int rc = check_polyphase_pass(c, w); if (rc) return rc;
    if (w.pf_state < 1 || PERF_FBMC_CW != 1) return CHEST_OK;
    const double removed = std::max(w.rsup_zeroed_max, std::max(s.mm[0].w_zeroed_max, s.mm[1].w_zeroed_max));
    // auto: only where it is also cheaper.  Measured break-even on B200: the modem chain costs 30-50 ns per column, the tile form
    // 8 P nnz flops at ~38 TFLOP/s -- CP-OFDM at the default geometry (1.0 MFLOP per column) is faster on the tiles.
    const double tile_flops = 8.0 * s.P * (double)std::max(s.mm[0].nnz_offdiag_pairs, s.mm[1].nnz_offdiag_pairs);
    on = c->est_mode == 2 || (removed < 1e-13 && (c->est_mode == 3 || tile_flops >= 3e6));
    return CHEST_OK;
}

// Phase B of iteration it for the factored schemes: H-hat of every column, then the modem chain with it.
int stage_factored_estimated_csi(Ctx* c, int n_rep, int it, int n_iter, const IcParams& ip) {
    const int N = c->N, var_prev = (it - 1 == 0 || (it - 1) <= n_iter / 2) ? 0 : 1;      // the D-hat estimated in iteration it-1 (DS.m:475,492)
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.ef_n_units) continue;
        EstChanParams ep{};
        ep.desc = w.ef_desc.p; ep.Mq = w.Mq.p; ep.hest = c->hest.p; ep.n_rep = n_rep; ep.TN = c->T * N;
        int Pmax = 1;
        for (int si = 0; si < 3; ++si) {
            Scheme& s = c->sch[si];
            if (!s.set || !c->est_fact[si]) continue;
            ep.hP[si] = ip.sch[si].hP; ep.rinv[si] = s.mm[var_prev].rinv.p; ep.P[si] = s.P; Pmax = std::max(Pmax, s.P);
        }
        const size_t smem_c = (size_t)2 * Pmax * NC_MAX * sizeof(cplx);
        if (smem_c > 48 * 1024) return fail(CHEST_ERR_STATE, "factored estimator: more than 96 pilots per scheme");
        const int P4 = (Pmax + 3) / 4;
        const size_t smem_m = (size_t)(Pmax + 4 * P4) * NC_MAX * sizeof(cplx);
        const bool scalar = getenv("CHEST_EST_CHANNEL_SCALAR") != nullptr;     // development / tests: the DFMA form
        bool same_p = true;                                                    // the DMMA form is instantiated per pilot-quad count
        for (int si = 0; si < 3; ++si) if (c->sch[si].set && c->est_fact[si] && c->sch[si].waveform == wfi && (c->sch[si].P + 3) / 4 != P4) same_p = false;
        if (!scalar && same_p && P4 == 4) k_est_channel_mma<4><<<w.ef_n_units, EST_CHAN_THREADS, smem_m, c->stream>>>(ep);
        else if (!scalar && same_p && P4 == 2) k_est_channel_mma<2><<<w.ef_n_units, EST_CHAN_THREADS, smem_m, c->stream>>>(ep);
        else k_est_channel<<<w.ef_n_units, EST_CHAN_THREADS, smem_c, c->stream>>>(ep);
        c->launches++;
        CK(cudaGetLastError());
        const ModemDev& md = w.modem;
        EstFactParams fp{};
        fp.md = md; fp.desc = w.ef_desc.p; fp.K_max = c->K_max; fp.n_rep = n_rep; fp.T = c->T; fp.N = N; fp.K = w.K;
        for (int si = 0; si < 3; ++si) fp.y[si] = ip.sch[si].y;
        fp.hest = c->hest.p; fp.tap_delay = c->d_tap_delay.p; fp.scratch = c->scratch.p; fp.colmajor = c->chain_cm ? 1 : 0;
        const size_t nbuf = std::max((size_t)md.Ksym * md.nfft, (size_t)N);
        const size_t smem = (2 * nbuf + md.nfft) * sizeof(cplx) + (size_t)md.Np * sizeof(double) + (size_t)md.L * sizeof(int);
        const bool no24 = getenv("CHEST_NO_FAST24") != nullptr;
        if (modem_fast24(md) && !no24 && w.K <= CHAIN24_NE * PERF_FBMC_THREADS && N <= CHAIN24_NH * PERF_FBMC_THREADS) k_est_factored<true><<<w.ef_n_units * NC_MAX, PERF_FBMC_THREADS, smem, c->stream>>>(fp);
        else k_est_factored<false><<<w.ef_n_units * NC_MAX, PERF_FBMC_THREADS, smem, c->stream>>>(fp);
        c->launches++;
        CK(cudaGetLastError());
    }
    return CHEST_OK;
}

int stage_factored_perfect_csi(Ctx* c, int n_rep, int it, int n_iter, uint32_t* err, const IcParams& ip) {
    const int N = c->N, Np = (N + 1) & ~1;
    static const bool legacy = getenv("CHEST_CHAIN_LEGACY") != nullptr;      // development: the block-barrier k_gemm<PLAIN> chain
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.set || !w.nsch || !w.f_cols) continue;
        if (!legacy && w.det_on) {                                 // FBMC: polyphase modem + equalisation + detection + counters
            const ModemDev& md = w.modem;
            PerfDetParams dp{};
            dp.md = md; dp.it = it; dp.n_iter = n_iter; dp.n_snr = c->S; dp.n_rep = n_rep; dp.nsch = w.nsch; dp.n_cols = w.f_cols;
            dp.T = c->T; dp.N = N; dp.K = w.K; dp.zw_stride = w.zw_stride;
            for (int q = 0; q < w.nsch; ++q) { dp.sch[q] = ip.sch[w.sch[q]]; dp.scheme_id[q] = w.sch[q]; }
            for (int k = 0; k < 2; ++k) dp.cst[k] = c->cst[k].dev;
            dp.voff = w.f_voff.p; dp.yoff = w.f_yoff.p; dp.rep = w.f_rep.p; dp.v_base = c->scratch.p;
            dp.y = w.y.p; dp.htrue = w.htrue.p; dp.h = c->h.p; dp.tap_delay = c->d_tap_delay.p; dp.zw_g = w.zw_g.p; dp.err = err;
            dp.vstride = c->chain_cm_perf ? 1 : NC_MAX;
            const size_t nbuf = std::max((size_t)md.Ksym * md.nfft, (size_t)N);
            const size_t smem = (2 * nbuf + md.nfft) * sizeof(cplx) + (size_t)md.Np * sizeof(double) + (size_t)md.L * sizeof(int);
            const bool no24 = getenv("CHEST_NO_FAST24") != nullptr;             // development / tests: the generic chain
            if (modem_fast24(md) && !no24 && w.K <= CHAIN24_NE * PERF_FBMC_THREADS && N <= CHAIN24_NH * PERF_FBMC_THREADS) k_perfect_fbmc_det<true><<<w.f_cols, PERF_FBMC_THREADS, smem + CHAIN24_NE * PERF_FBMC_THREADS * sizeof(cplx) + 16, c->stream>>>(dp);
            else k_perfect_fbmc_det<false><<<w.f_cols, PERF_FBMC_THREADS, smem, c->stream>>>(dp);
            c->launches++;
            CK(cudaGetLastError());
            continue;
        }
        if (it == 0) continue;                                      // (the one-tap stage of the other columns runs in k_ic_light)
        if (!legacy && w.pf_state == 1 && w.pf_n_groups > 0) {                             // FBMC: the polyphase modem instead of the GEMMs
            const ModemDev& md = w.modem;
            PerfFbmcParams pp{};
            pp.md = md; pp.n_groups = w.pf_n_groups; pp.T = c->T; pp.N = N; pp.K = w.K; pp.groups = w.pf_groups.p;
            pp.voff = w.f_voff.p; pp.yoff = w.f_yoff.p; pp.rep = w.f_rep.p;
            pp.v_base = c->scratch.p; pp.y_base = c->scratch.p + (size_t)c->K_max * NC_MAX;
            pp.y = w.y.p; pp.htrue = w.htrue.p; pp.h = c->h.p; pp.tap_delay = c->d_tap_delay.p;
            const size_t smem = ((size_t)2 * PERF_FBMC_CW * md.Ksym * md.nfft + md.nfft) * sizeof(cplx) + (size_t)md.Np * sizeof(double) + (size_t)md.L * sizeof(int);
            k_perfect_fbmc<PERF_FBMC_CW><<<w.pf_n_groups, PERF_FBMC_THREADS, smem, c->stream>>>(pp);
            c->launches++;
            CK(cudaGetLastError());
            continue;
        }
        if (!legacy) {
            GemmRingParams p{};                                                             // s = G v
            p.M = N; p.n_cols = w.f_cols; p.lda = (w.K + 1) & ~1; p.ldc = N;
            p.At1 = w.Gt1.p; p.At2 = w.Gt2.p; p.mt_klo = w.gt_klo.p; p.mt_khi = w.gt_khi.p; p.m8_klo = w.gt8_klo.p; p.m8_khi = w.gt8_khi.p;
            p.bsrc = c->scratch.p; p.b_off = w.f_voff.p; p.b_kstride = NC_MAX; p.out = w.f_s.p;
            CK((launch_gemm_ring<true, 0>(c, p, w.tile)));
            dim3 grid(w.f_cols, (N + 127) / 128);                                           // r = H s, written as operand planes
            k_apply_h_cols_planes<<<grid, 128, 0, c->stream>>>(w.f_r1.p, w.f_r2.p, w.f_s.p, c->h.p, c->d_tap_delay.p, w.f_rep.p, N, Np, c->T);
            c->launches++;
            CK(cudaGetLastError());
            GemmRingParams q{};                                                             // y_ic = y - Q^H r + h v
            q.M = w.K; q.n_cols = w.f_cols; q.lda = Np; q.ldb = Np; q.ldc = w.K;
            q.At1 = w.Q1.p; q.At2 = w.Q2.p; q.mt_klo = w.q_klo.p; q.mt_khi = w.q_khi.p; q.m8_klo = w.q8_klo.p; q.m8_khi = w.q8_khi.p;
            q.b1 = w.f_r1.p; q.b2 = w.f_r2.p;
            q.e_out = c->scratch.p + (size_t)c->K_max * NC_MAX; q.e_off = w.f_voff.p;       // y_ic sits one buffer behind v
            q.e_y = w.y.p; q.e_yoff = w.f_yoff.p; q.e_h = w.htrue.p; q.e_rep = w.f_rep.p; q.e_v = c->scratch.p;
            CK((launch_gemm_ring<false, 1>(c, q, w.tile)));
            continue;
        }
        GemmParams p{};
        p.M = N; p.Kc = w.K; p.n_cols = w.f_cols; p.lda = w.K; p.ldc = N; p.conj_a = 0;
        p.At = w.Gt.p; p.mt_klo = w.gt_klo.p; p.mt_khi = w.gt_khi.p; p.out = w.f_s.p;
        p.m8_klo = w.gt8_klo.p; p.m8_khi = w.gt8_khi.p;
        p.bsrc = c->scratch.p; p.b_off = w.f_voff.p; p.b_kstride = NC_MAX;
        CK(launch_gemm<GEMM_PLAIN>(c, p, 1, w.tile));                                   // s = G v
        dim3 grid(w.f_cols, (N + 127) / 128);
        k_apply_h_cols<<<grid, 128, 0, c->stream>>>(w.f_r.p, w.f_s.p, c->h.p, c->d_tap_delay.p, w.f_rep.p, N, c->T);   // r = H s
        c->launches++;
        CK(cudaGetLastError());
        GemmParams q{};
        q.M = w.K; q.Kc = N; q.n_cols = w.f_cols; q.lda = N; q.ldc = w.K; q.conj_a = 1;
        q.At = w.Q.p; q.mt_klo = w.q_klo.p; q.mt_khi = w.q_khi.p; q.out = w.f_s.p;
        q.m8_klo = w.q8_klo.p; q.m8_khi = w.q8_khi.p;
        q.bsrc = w.f_r.p; q.ldb = N;
        q.e_out = c->scratch.p + (size_t)c->K_max * NC_MAX; q.e_off = w.f_voff.p;       // y_ic sits one buffer behind v
        q.e_y = w.y.p; q.e_yoff = w.f_yoff.p; q.e_h = w.htrue.p; q.e_rep = w.f_rep.p;
        q.e_v = c->scratch.p; q.e_voff = w.f_voff.p;
        CK(launch_gemm<GEMM_PLAIN>(c, q, 1, w.tile));                                   // y_ic = y - Q^H r + h v
    }
    return CHEST_OK;
}

// ---- split-BF16 tensor-core mode: operand images of W (kernels_tc.cuh).  Once per (scheme, variant); all SNR points.
int ensure_tc_pack(Ctx* c) {
    int p8 = 16;
    for (int si = 0; si < 3; ++si) if (c->sch[si].set) p8 = std::max(p8, c->sch[si].P <= 16 ? 16 : 32);
    for (int si = 0; si < 3; ++si)
        if (c->sch[si].set && c->sch[si].P > 32) return fail(CHEST_ERR_STATE, "split-BF16 mode supports at most 32 pilots per scheme");
    if (c->tc_p8 != p8) { for (auto& s : c->sch) for (auto& m : s.mm) m.tc_packed = false; c->tc_p8 = p8; }
    const int NCH = p8 / 4;
    const size_t a_bytes = (size_t)2 * NCH * TC_ROWS * 16;
    cudaStream_t st = c->stream;
    for (int si = 0; si < 3; ++si) {
        Scheme& s = c->sch[si];
        if (!s.set) continue;
        const int K = s.K, P4 = (s.P + 3) / 4, RT8 = (K + 7) / 8, n_rt = (K + TC_ROWS - 1) / TC_ROWS;
        for (int v = 0; v < 2; ++v) {
            MmseVariant& m = s.mm[v];
            if (!m.set || m.tc_packed) continue;
            std::vector<int> tptr(RT8 + 1), tdel(std::max(m.n_tiles, 1));
            CK(cudaMemcpyAsync(tptr.data(), m.tile_ptr.p, sizeof(int) * (RT8 + 1), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(tdel.data(), m.tile_delta.p, sizeof(int) * std::max(m.n_tiles, 1), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            std::vector<int> jlist, jptr(n_rt + 1, 0), a_tile;
            std::vector<char> mark(K);
            for (int t = 0; t < n_rt; ++t) {
                std::fill(mark.begin(), mark.end(), 0);
                for (int rt = t * (TC_ROWS / 8); rt < std::min(RT8, (t + 1) * (TC_ROWS / 8)); ++rt)
                    for (int e = tptr[rt]; e < tptr[rt + 1]; ++e)
                        for (int r = 0; r < 8; ++r) {
                            const int i = rt * 8 + r, j = i + tdel[e];
                            if (i < K && j >= 0 && j < K) mark[j] = 1;
                        }
                for (int j = 0; j < K; ++j) if (mark[j]) { jlist.push_back(j); a_tile.push_back(t); }
                jptr[t + 1] = (int)jlist.size();
            }
            m.tc_entries = (int)jlist.size(); m.tc_row_tiles = n_rt; m.tc_jptr_h = jptr;
            if (jlist.empty()) { jlist.push_back(0); a_tile.push_back(0); }
            DevBuf<int> d_atile;
            CK(m.tc_jlist.upload(jlist, st)); CK(m.tc_jptr.upload(jptr, st)); CK(d_atile.upload(a_tile, st));
            CK(m.tc_img.alloc((size_t)c->S * std::max(m.tc_entries, 1) * a_bytes));
            for (int snr = 0; snr < c->S && m.tc_entries > 0; ++snr) {
                k_tc_pack_w<<<m.tc_entries, TC_ROWS, 0, st>>>(m.tc_img.p + (size_t)snr * m.tc_entries * a_bytes, m.frag[snr].p, m.tile_ptr.p,
                                                              m.tile_delta.p, m.tc_jlist.p, d_atile.p, K, P4, NCH);
                c->launches++;
            }
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));
            m.tc_packed = true;
        }
    }
    return CHEST_OK;
}

// Work items of k_ic_est_tc: (scheme, SNR point, up to 8 consecutive EST units = 128 realization columns, one row tile),
// balanced over one CTA per SM on the host (longest processing time first).
int build_tc_items(Ctx* c, const std::vector<IcCta>& v, int n_est) {
    struct Cost { TcItem it; int cost; };
    std::vector<Cost> all;
    const int NCH = c->tc_p8 / 4;
    double flops = 0;
    for (int u = 0; u < n_est;) {
        int u1 = u + 1;
        while (u1 < n_est && u1 - u < TC_COLS / NC_MAX && v[u1].mode == 0 && v[u1].scheme_or_wf == v[u].scheme_or_wf && v[u1].snr == v[u].snr) ++u1;
        const MmseVariant& m = c->sch[v[u].scheme_or_wf].mm[0];
        for (int t = 0; t < m.tc_row_tiles; ++t) {
            const int cost = m.tc_jptr_h[t + 1] - m.tc_jptr_h[t];
            all.push_back({{v[u].scheme_or_wf, v[u].snr, u, u1 - u, t, t + 1}, cost});
            flops += (double)cost * 3 * (NCH / 2) * 2.0 * TC_ROWS * (2 * TC_COLS) * 16;
        }
        u = u1;
    }
    c->tc_mma_flops = flops;
    std::stable_sort(all.begin(), all.end(), [](const Cost& a, const Cost& b) { return a.cost > b.cost; });
    const int grid = std::max(1, std::min(c->n_sm, (int)all.size()));
    std::vector<std::vector<TcItem>> per(grid);
    std::vector<long long> load(grid, 0);
    // longest first onto the least loaded CTA (a heap is not worth it for a few thousand items x 148 CTAs)
    for (const Cost& x : all) {
        int best = 0;
        for (int b = 1; b < grid; ++b) if (load[b] < load[best]) best = b;
        per[best].push_back(x.it); load[best] += x.cost + 8;
    }
    std::vector<TcItem> items; std::vector<int> ptr(grid + 1, 0);
    for (int b = 0; b < grid; ++b) { items.insert(items.end(), per[b].begin(), per[b].end()); ptr[b + 1] = (int)items.size(); }
    c->tc_n_items = (int)items.size(); c->tc_grid = grid;
    if (items.empty()) items.push_back({0, 0, 0, 0, 0, 0});
    CK(c->tc_items.upload(items, c->stream)); CK(c->tc_cta_ptr.upload(ptr, c->stream));
    CK(c->tc_status.alloc(1)); CK(cudaMemsetAsync(c->tc_status.p, 0, sizeof(int), c->stream));
    return CHEST_OK;
}

int build_ctas(Ctx* c, int n_rep) {
    if (c->ctas_for_batch == n_rep) return CHEST_OK;
    std::vector<IcCta> v;
    auto est = [&](int si) {
        if (!c->sch[si].set) return;
        for (int snr = 0; snr < c->S; ++snr)
            for (int b0 = 0; b0 < n_rep; b0 += NC_MAX) v.push_back({0, si, snr, b0, std::min(NC_MAX, n_rep - b0)});
    };
    auto perf = [&](int wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.set || w.nsch == 0) return;
        // columns: 8 * (scheme slot) + (SNR point - first); one unit per block of 8 SNR points
        for (int b = 0; b < n_rep; ++b)
            for (int s0 = 0; s0 < c->S; s0 += 8)
                v.push_back({1, wfi, b, s0, w.nsch == 2 ? 16 : std::min(8, c->S - s0)});
    };
    // heavy work first; compute-bound EST CTAs and memory-heavy PERF CTAs are interleaved so that both the
    // FP64 tensor pipe and HBM stay busy
    std::vector<IcCta> e1, p1, e2, p2;
    for (int wfi = 0; wfi < 2; ++wfi) {                        // waveforms whose perfect-CSI twin runs as one fused kernel
        Waveform& w = c->wf[wfi];
        w.twin_on = false;
        // measured on B200 (default geometry, B = 4096): 20.9 ms per step against 15.8 ms for PERF units + k_perfect_fbmc -- the
        // long precoder rows (64 auxiliary rows x 204 entries) are re-read per column here, while k_ic_light shares them
        // across 16 columns on DMMA tiles -- so the fused twin stays opt-in (CHEST_TWIN=1); parity-tested either way
        if (!w.set || !w.nsch || c->perf_mode != 1 || !getenv("CHEST_TWIN")) continue;
        int rc = check_polyphase_pass(c, w); if (rc) return rc;
        bool ok = w.pf_state == 1;
        for (int q = 0; q < w.nsch && ok; ++q) ok = c->cst[c->sch[w.sch[q]].constellation].order <= 256;
        w.twin_on = ok;
    }
    est(CHEST_SCHEME_AUX); est(CHEST_SCHEME_COD); e1.swap(v);
    if (!c->wf[CHEST_WF_FBMC].twin_on) perf(CHEST_WF_FBMC);
    p1.swap(v);
    est(CHEST_SCHEME_OFDM); e2.swap(v);
    if (!c->wf[CHEST_WF_OFDM].twin_on) perf(CHEST_WF_OFDM);
    p2.swap(v);
    auto weave = [&](std::vector<IcCta>& a, std::vector<IcCta>& b) {
        size_t ia = 0, ib = 0;
        while (ia < a.size() || ib < b.size()) {
            // keep the ratio a:b constant along the list
            if (ib >= b.size() || (ia < a.size() && ia * b.size() <= ib * a.size())) v.push_back(a[ia++]);
            else v.push_back(b[ib++]);
        }
    };
    if (const char* dbg = getenv("CHEST_DEBUG_ONLY")) {      // development knob: time one CTA class alone
        if (!strcmp(dbg, "est")) { p1.clear(); p2.clear(); }
        if (!strcmp(dbg, "perf")) { e1.clear(); e2.clear(); }
        if (!strcmp(dbg, "est_fbmc")) { p1.clear(); p2.clear(); e2.clear(); }
        if (!strcmp(dbg, "perf_fbmc")) { e1.clear(); e2.clear(); p2.clear(); }
    }
    for (int si = 0; si < 3; ++si) c->est_fact[si] = false;
    for (auto& w : c->wf) w.ef_n_units = 0;
    c->chain_cm = getenv("CHEST_CHAIN_ROWMAJOR") == nullptr;              // development knobs: the row-major layout everywhere / for the PERF units
    c->chain_cm_perf = c->chain_cm && getenv("CHEST_PERF_ROWMAJOR") == nullptr;
    if (c->perf_mode == 1) {                                   // factored mode: k_ic_main runs the EST units only (listed first)
        // schemes whose estimated-CSI cancellation runs in factored form (k_est_channel + k_est_factored) come after the tile-form
        // EST units: k_ic_main stops at n_est_units, k_ic_light walks all of them
        std::vector<IcCta> ef;
        std::vector<int4> ef_list[2];
        for (int si = 0; si < 3; ++si) {
            int rc = est_factored_usable(c, si, c->est_fact[si]); if (rc) return rc;
        }
        for (auto* lst : {&e1, &e2})
            for (const IcCta& u : *lst) (c->est_fact[u.scheme_or_wf] ? ef : v).push_back(u);
        c->n_est_units = (int)v.size();
        for (const IcCta& u : ef) { ef_list[c->sch[u.scheme_or_wf].waveform].push_back(make_int4((int)v.size(), u.scheme_or_wf, u.snr, u.first)); v.push_back(u); }
        size_t ef_cols = 0;
        for (int wfi = 0; wfi < 2; ++wfi) {
            Waveform& w = c->wf[wfi];
            w.ef_n_units = (int)ef_list[wfi].size();
            if (w.ef_n_units) CK(w.ef_desc.upload(ef_list[wfi], c->stream));
            ef_cols = std::max(ef_cols, (size_t)w.ef_n_units * NC_MAX);
        }
        if (ef_cols) CK(c->hest.alloc(ef_cols * c->T * c->N));
        c->wf[CHEST_WF_FBMC].perf_base = (int)v.size(); v.insert(v.end(), p1.begin(), p1.end());
        c->wf[CHEST_WF_OFDM].perf_base = (int)v.size(); v.insert(v.end(), p2.begin(), p2.end());
    } else if (getenv("CHEST_IC_WEAVE")) {                     // development: interleave compute- and memory-heavy units
        weave(e1, p1);
        weave(e2, p2);
    } else {                                                   // longest units first: the short ones fill the tail of the queue
        v.insert(v.end(), e1.begin(), e1.end()); v.insert(v.end(), p1.begin(), p1.end());
        v.insert(v.end(), e2.begin(), e2.end()); v.insert(v.end(), p2.begin(), p2.end());
    }
    c->n_ctas = (int)v.size();
    if (c->perf_mode != 1) c->n_est_units = c->n_ctas;
    if (c->precision == 1) {
        if (c->perf_mode != 1) return fail(CHEST_ERR_STATE, "the split-BF16 mode runs with the factored perfect-CSI pass (CHEST_PERFECT_FACTORED)");
        int rc = ensure_tc_pack(c); if (rc) return rc;
        if (c->n_est_units > 0) { rc = build_tc_items(c, v, c->n_est_units); if (rc) return rc; }
    }
    CK(c->ctas.upload(v, c->stream));
    CK(c->scratch.alloc((size_t)c->n_ctas * 3 * c->K_max * NC_MAX));
    if (c->perf_mode == 1) {
        // column tables: perfect-CSI column (rep, scheme slot, snr) -> its v / y_ic slot in the unit scratch, its y
        const int S = c->S, nblk = (S + 7) / 8;
        for (int wfi = 0; wfi < 2; ++wfi) {
            Waveform& w = c->wf[wfi];
            w.det_on = false;
            if (!w.set || !w.nsch || w.twin_on) { w.f_cols = 0; w.pf_n_groups = 0; continue; }
            const int nv = w.nsch * S;
            w.perf_nblk = nblk; w.f_cols = nv * n_rep;
            {   // one column per CTA with detection behind the chain (k_perfect_fbmc_det)?  Decides the layout of these units' v
                int rc = check_polyphase_pass(c, w); if (rc) return rc;
                const char* le = getenv("CHEST_LIGHT");
                w.det_on = w.pf_state >= 1 && PERF_FBMC_CW == 1 && !getenv("CHEST_NO_PERF_DETECT") && !getenv("CHEST_CHAIN_LEGACY") && !(le && !strcmp(le, "post"));
            }
            const bool vcm = w.det_on && c->chain_cm_perf;
            std::vector<int64_t> voff(w.f_cols), yoff(w.f_cols);
            std::vector<int> rep(w.f_cols);
            for (int r = 0; r < n_rep; ++r)
                for (int slot = 0; slot < w.nsch; ++slot)
                    for (int snr = 0; snr < S; ++snr) {
                        const int col = r * nv + slot * S + snr;
                        const int64_t unit = w.perf_base + (int64_t)r * nblk + snr / 8;
                        voff[col] = (unit * 3 + 1) * c->K_max * NC_MAX + (vcm ? (int64_t)(slot * 8 + snr % 8) * c->K_max : slot * 8 + snr % 8);
                        yoff[col] = ((int64_t)slot * S * n_rep + (int64_t)snr * n_rep + r) * w.K;
                        rep[col] = r;
                    }
            CK(w.f_voff.upload(voff, c->stream)); CK(w.f_yoff.upload(yoff, c->stream)); CK(w.f_rep.upload(rep, c->stream));
            {   // column groups of the polyphase pass: up to CW columns that are neighbours in one unit's scratch
                int rc = check_polyphase_pass(c, w); if (rc) return rc;
                std::vector<int2> groups;
                if (w.pf_state == 1)
                    for (int r = 0; r < n_rep; ++r)
                        for (int slot = 0; slot < w.nsch; ++slot)
                            for (int s0 = 0; s0 < S; s0 += 8)
                                for (int q = s0; q < std::min(S, s0 + 8); q += PERF_FBMC_CW)
                                    groups.push_back(make_int2(r * nv + slot * S + q, std::min(PERF_FBMC_CW, std::min(S, s0 + 8) - q)));
                w.pf_n_groups = (int)groups.size();
                if (!groups.empty()) CK(w.pf_groups.upload(groups, c->stream));
                // one column per CTA: the same kernel also equalises, detects and counts (k_perfect_fbmc_det); k_ic_light then
                // only precodes these columns.  CHEST_NO_PERF_DETECT keeps the two-kernel split (development / tests).
                if (w.det_on) {
                    int nd = 0;
                    for (int q = 0; q < w.nsch; ++q) nd = std::max(nd, c->sch[w.sch[q]].n_data);
                    w.zw_stride = (nd + 15) & ~15;
                    CK(w.zw_g.alloc((size_t)w.f_cols * w.zw_stride));
                }
            }
            CK(w.f_s.alloc((size_t)w.f_cols * c->N)); CK(w.f_r.alloc((size_t)w.f_cols * c->N));
            {
                const size_t Np = (size_t)((c->N + 1) & ~1);
                const bool fresh = w.f_r1.n < (size_t)w.f_cols * Np;
                CK(w.f_r1.alloc((size_t)w.f_cols * Np)); CK(w.f_r2.alloc((size_t)w.f_cols * Np));
                if (fresh) {                                    // the padding sample of an odd N is read by the 16-byte copies
                    CK(cudaMemsetAsync(w.f_r1.p, 0, sizeof(cplx) * w.f_r1.n, c->stream));
                    CK(cudaMemsetAsync(w.f_r2.p, 0, sizeof(double) * w.f_r2.n, c->stream));
                }
            }
        }
    }
    CK(cudaStreamSynchronize(c->stream));
    c->ctas_for_batch = n_rep;
    if (getenv("CHEST_VERBOSE"))
        fprintf(stderr, "chest: units %d (tile-form EST %d), polyphase state FBMC %d OFDM %d, detect-in-chain %d %d, factored estimator aux %d cod %d ofdm %d\n",
                c->n_ctas, c->n_est_units, c->wf[0].pf_state, c->wf[1].pf_state, (int)c->wf[0].det_on, (int)c->wf[1].det_on,
                (int)c->est_fact[0], (int)c->est_fact[1], (int)c->est_fact[2]);
    return CHEST_OK;
}

int finish_pipeline(Ctx* c);

// Enqueues the whole loop body on c->stream.  wait = false leaves the run pending (chest_wait collects it): nothing
// in here blocks the host once the per-batch-size tables exist, so one host thread can drive several devices.
int run_pipeline(Ctx* c, int n_rep, int n_iter, const chest_draws* draws, uint64_t seed, int64_t first_rep,
                 uint32_t* err_host, uint32_t* err_dev, bool wait = true) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (c->pending) return fail(CHEST_ERR_STATE, "an asynchronous run is pending on this context: call chest_wait first");
    ARG(n_rep >= 1 && n_rep <= c->max_batch);
    ARG(n_iter >= 0 && n_iter <= 16);
    ARG(c->S >= 1);
    for (int si = 0; si < 3; ++si)
        if (c->sch[si].set && (!c->sch[si].mm[0].set || !c->sch[si].mm[1].set)) return fail(CHEST_ERR_STATE, "scheme without both MMSE variants");
    if (!(c->fD > 0)) return fail(CHEST_ERR_STATE, "the batched loop body synthesises time-variant realizations: f_D must be positive");
    const int S = c->S, N = c->N, TP = c->T * c->paths;
    c->cur_batch = n_rep; c->last_iter = n_iter;
    cudaStream_t st = c->stream;
    if (c->profiling) CK(cudaEventRecord(c->ev[0], st));
    // ---- stage 0: draws
    const double *du = c->doppler_u.p, *pu = c->phase_u.p;
    const cplx* noise = c->noise.p;
    const uint8_t* bits[3] = {c->sch[0].bits.p, c->sch[1].bits.p, c->sch[2].bits.p};
    const int32_t* pidx[2] = {c->pilot_idx[0].p, c->pilot_idx[1].p};
    int pf_used = -1;
    const bool discrete = c->n_shift > 0;
    const cplx* cgauss = c->chan_gauss.p;
    const size_t n_cg = (size_t)(2 * c->n_shift + 1) * c->T;
    if (draws && discrete) {                                   // the discrete spectrum draws normals, not uniforms (FF.m:208-209)
        ARG(draws->channel_gauss && draws->noise);
        if (draws->on_device) {
            for (int q = 0; q < 2; ++q)
                if (c->pf[q].used && draws->channel_gauss == reinterpret_cast<const double*>(c->pf[q].cg.p)) { pf_used = q; CK(cudaStreamWaitEvent(st, c->pf[q].landed, 0)); }
            cgauss = reinterpret_cast<const cplx*>(draws->channel_gauss);
        } else CK(cudaMemcpyAsync(c->chan_gauss.p, draws->channel_gauss, sizeof(cplx) * n_rep * n_cg, cudaMemcpyHostToDevice, st));
    }
    if (draws && draws->on_device && discrete) {
        for (int i = 0; i < 3; ++i) if (c->sch[i].set) ARG(draws->bits[i]);
        for (int i = 0; i < 2; ++i) if (c->wf[i].set && c->wf[i].nsch) ARG(draws->pilot_idx[i]);
        noise = reinterpret_cast<const cplx*>(draws->noise);
        for (int i = 0; i < 3; ++i) bits[i] = draws->bits[i];
        for (int i = 0; i < 2; ++i) pidx[i] = draws->pilot_idx[i];
    } else if (draws && draws->on_device) {
        ARG(draws->doppler_u && draws->phase_u && draws->noise);
        for (int i = 0; i < 3; ++i) if (c->sch[i].set) ARG(draws->bits[i]);
        for (int i = 0; i < 2; ++i) if (c->wf[i].set && c->wf[i].nsch) ARG(draws->pilot_idx[i]);
        for (int q = 0; q < 2; ++q)
            if (c->pf[q].used && draws->doppler_u == c->pf[q].du.p) { pf_used = q; CK(cudaStreamWaitEvent(st, c->pf[q].landed, 0)); }
        du = draws->doppler_u; pu = draws->phase_u; noise = reinterpret_cast<const cplx*>(draws->noise);
        for (int i = 0; i < 3; ++i) bits[i] = draws->bits[i];
        for (int i = 0; i < 2; ++i) pidx[i] = draws->pilot_idx[i];
    } else if (draws) {
        ARG(draws->noise && (discrete || (draws->doppler_u && draws->phase_u)));
        if (!discrete) {
            CK(cudaMemcpyAsync(c->doppler_u.p, draws->doppler_u, sizeof(double) * n_rep * TP, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(c->phase_u.p, draws->phase_u, sizeof(double) * n_rep * TP, cudaMemcpyHostToDevice, st));
        }
        CK(cudaMemcpyAsync(c->noise.p, draws->noise, sizeof(cplx) * (size_t)n_rep * S * N, cudaMemcpyHostToDevice, st));
        for (int i = 0; i < 3; ++i)
            if (c->sch[i].set) {
                ARG(draws->bits[i]);
                CK(cudaMemcpyAsync(c->sch[i].bits.p, draws->bits[i], (size_t)n_rep * c->sch[i].n_bits, cudaMemcpyHostToDevice, st));
            }
        for (int i = 0; i < 2; ++i)
            if (c->wf[i].set && c->wf[i].nsch) {
                ARG(draws->pilot_idx[i]);
                int P = c->sch[c->wf[i].sch[0]].P;
                CK(cudaMemcpyAsync(c->pilot_idx[i].p, draws->pilot_idx[i], sizeof(int32_t) * n_rep * P, cudaMemcpyHostToDevice, st));
            }
    } else {
        chest_draws tmp;
        rc = chest_generate_draws((uint64_t)(uintptr_t)c, n_rep, seed, first_rep, &tmp);
        if (rc) return rc;
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[1], st));
    // ---- stage 1 (K1): channel realization, TX symbols, s = G x, r0 = H s
    rc = discrete ? stage_channel_discrete(c, n_rep, cgauss) : stage_channel(c, n_rep, du, pu);
    if (rc) return rc;
    if (c->profiling) CK(cudaEventRecord(c->ev_k1[0], st));
    for (int si = 0; si < 3; ++si) {
        if (!c->sch[si].set) continue;
        SchemeDev sd = scheme_dev(c, si);
        k_tx_symbols<<<n_rep, 256, sd.K_in * sizeof(cplx), st>>>(sd, c->cst[sd.constellation].dev, bits[si],
                                                                 pidx[sd.waveform], n_rep);
        c->launches++;
        CK(cudaGetLastError());
    }
    if (c->profiling) CK(cudaEventRecord(c->ev_k1[1], st));
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (c->profiling) CK(cudaEventRecord(c->ev_k1[2 + 3 * wfi], st));
        if (w.set && w.nsch) {
            GemmParams p{};
            p.M = N; p.Kc = w.K; p.n_cols = w.nsch * n_rep; p.lda = w.K; p.ldc = N; p.conj_a = 0;
            p.At = w.Gt.p; p.mt_klo = w.gt_klo.p; p.mt_khi = w.gt_khi.p; p.out = w.s.p;
            p.bsrc = w.x.p; p.ldb = w.K;
            CK(launch_gemm<GEMM_PLAIN>(c, p, 1, w.tile));
        }
        if (c->profiling) CK(cudaEventRecord(c->ev_k1[3 + 3 * wfi], st));
        if (w.set && w.nsch) {
            dim3 grid((N + 127) / 128, w.nsch * n_rep);
            k_apply_h<<<grid, 128, 0, st>>>(w.r0.p, w.s.p, c->h.p, c->d_tap_delay.p, N, c->T, n_rep, -1);
            c->launches++;
            CK(cudaGetLastError());
        }
        if (c->profiling) CK(cudaEventRecord(c->ev_k1[4 + 3 * wfi], st));
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[2], st));
    // ---- stage 2 (K2): D = Q^H H G, h = diag(D)
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.set || !w.nsch) continue;
        if (c->perf_mode == 1) {                               // factored mode needs only h = diag(D)
            rc = ensure_pd(c, w); if (rc) return rc;
            GemmParams p{};                                    // htrue[rep][i] = P[i][:] . h[rep][:]
            p.M = w.K; p.Kc = c->T * N; p.n_cols = n_rep; p.lda = c->T * N; p.ldc = w.K; p.conj_a = 0;
            p.At = w.Pd.p; p.mt_klo = w.pd_klo.p; p.mt_khi = w.pd_khi.p; p.out = w.htrue.p;
            p.bsrc = c->h.p; p.ldb = c->T * N;
            CK(launch_gemm<GEMM_PLAIN>(c, p, 1, w.tile));
        } else { rc = stage_transmission_matrix(c, wfi, n_rep, 0); if (rc) return rc; }
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[3], st));
    // ---- stage 3 (K3a): y = Q^H (r0 + noise) for every (scheme, SNR, realization)
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.set || !w.nsch) continue;
        { rc = check_polyphase_pass(c, w); if (rc) return rc; }
        if (w.pf_state >= 1 && !getenv("CHEST_CHAIN_LEGACY")) {          // the FFT-form demodulator instead of the GEMM
            DemodFbmcParams dp{};
            dp.md = w.modem; dp.N = N; dp.K = w.K; dp.n_snr = S; dp.n_rep = n_rep; dp.n_cols = w.nsch * S * n_rep;
            dp.r0 = w.r0.p; dp.noise = noise; dp.noise_scale = c->d_noise_scale.p; dp.y = w.y.p; dp.fast24 = getenv("CHEST_NO_FAST24") ? 0 : 1;
            const ModemDev& md = w.modem;
            const size_t nbuf = std::max((size_t)md.Ksym * md.nfft, (size_t)N);
            const size_t smem = (2 * nbuf + md.nfft) * sizeof(cplx) + (size_t)md.Np * sizeof(double) + (size_t)md.L * sizeof(int);
            k_demod_fbmc<<<dp.n_cols, PERF_FBMC_THREADS, smem, st>>>(dp);
            c->launches++;
            CK(cudaGetLastError());
            continue;
        }
        GemmParams p{};
        p.M = w.K; p.Kc = N; p.n_cols = w.nsch * S * n_rep; p.lda = (N + 1) & ~1; p.ldc = w.K; p.conj_a = 1;
        p.At1 = w.Q1.p; p.At2 = w.Q2.p; p.mt_klo = w.q_klo.p; p.mt_khi = w.q_khi.p; p.out = w.y.p;
        p.m8_klo = w.q8_klo.p; p.m8_khi = w.q8_khi.p;
        p.r0 = w.r0.p; p.noise = noise; p.noise_scale = c->d_noise_scale.p; p.n_snr = S; p.n_rep = n_rep; p.N = N;
        CK(launch_gemm<GEMM_DEMOD>(c, p, 1, w.tile));
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[4], st));
    // ---- stage 4/5 (K3b + K4): one-tap stage, then one fused launch per IC iteration
    rc = build_ctas(c, n_rep);
    if (rc) return rc;
    IcParams ip{};
    ip.n_iter = n_iter; ip.n_rep = n_rep; ip.n_snr = S; ip.K_max = c->K_max; ip.ctas = c->ctas.p;
    for (int si = 0; si < 3; ++si) if (c->sch[si].set) ip.sch[si] = scheme_dev(c, si);
    for (int k = 0; k < 2; ++k) ip.cst[k] = c->cst[k].dev;
    for (int wfi = 0; wfi < 2; ++wfi) {
        ip.wf_nscheme[wfi] = c->wf[wfi].nsch;
        ip.wf_scheme[wfi][0] = c->wf[wfi].sch[0]; ip.wf_scheme[wfi][1] = c->wf[wfi].sch[1];
        ip.D[wfi] = c->wf[wfi].D.p; ip.htrue[wfi] = c->wf[wfi].htrue.p;
        ip.d_jlo[wfi] = c->wf[wfi].d_jlo.p; ip.d_jhi[wfi] = c->wf[wfi].d_jhi.p;
    }
    ip.scratch = c->scratch.p;
    ip.pilot_rows = 4;
    for (int si = 0; si < 3; ++si) if (c->sch[si].set) ip.pilot_rows = std::max(ip.pilot_rows, 4 * ((c->sch[si].P + 3) / 4));
    size_t cst_smem = 0;                                           // shared-memory copies of the constellation tables
    for (int k = 0; k < 2; ++k)
        cst_smem += sizeof(cplx) * ((size_t)c->cst[k].order + (c->cst[k].n_axis + 1) / 2 + (c->cst[k].order + 3) / 4);
    const int ic_threads = IC_THREADS;
    ip.ring_cplx = (EST_WSRC == 0) ? EST_RING * (ip.pilot_rows / 4) * 32 : 0;   // EST_RING tiles of P4 fragments x 32 lanes
    // main: pilot tables (2) + max(v-chunk stages of a PERF unit, W-fragment rings of an EST unit)
    ip.stage_cplx = (int)((std::max((size_t)2 * PERF_STAGE_CPLX, (size_t)(ic_threads / 32) * ip.ring_cplx) + 7) / 8 * 8);
    const size_t main_smem = ((size_t)ip.stage_cplx + (size_t)2 * ip.pilot_rows * (NC_MAX + 2)) * sizeof(cplx)
                             + (size_t)ip.pilot_rows * EST_H1S * sizeof(double);
    // light / post: new pilot estimates, transmitted pilots, constellation tables, decided words (one byte per symbol);
    // k_ic_post adds the ring of y_ic chunks in front
    const size_t light_smem = (size_t)ip.pilot_rows * (2 * NC_MAX + 2) * sizeof(cplx) + cst_smem + (size_t)c->K_max * NC_MAX;
    const size_t post_smem = light_smem + (size_t)POST_NS * POST_ROWS * NC_MAX * sizeof(cplx);
    const char* light_env = getenv("CHEST_LIGHT");
    // k_ic_light is the default: on B200 it is the faster of the two (17.3 vs 22.1 ms per 4096-realization step,
    // profiles/r02_kic_post_vs_light.txt); CHEST_LIGHT=post selects the bulk-copy staged k_ic_post
    bool use_post = light_env && !strcmp(light_env, "post") && post_smem <= 110 * 1024;
    int p4_all = -1;                                              // pilot-quad count shared by every scheme (register prefetch)
    for (int si = 0; si < 3; ++si) if (c->sch[si].set) { int q = (c->sch[si].P + 3) / 4; p4_all = p4_all < 0 ? q : (p4_all == q ? q : 0); }
    void (*post_kernel)(IcParams) = p4_all == 4 ? k_ic_post<4> : (p4_all == 8 ? k_ic_post<8> : k_ic_post<0>);
    // which loop bodies the main kernel carries: the pilot-quad count most EST units have, and the real-v masks
    // every perfect-CSI unit of this configuration satisfies (16-column units: both schemes of a waveform)
    int p4s = 4, m2 = 3, m1 = 1;
    {
        int n4 = 0, n8 = 0;
        for (int si = 0; si < 3; ++si) if (c->sch[si].set) { int q = (c->sch[si].P + 3) / 4; n4 += q == 4; n8 += q == 8; }
        p4s = n8 > n4 ? 8 : 4;
        for (int wfi = 0; wfi < 2; ++wfi) {
            Waveform& w = c->wf[wfi];
            if (!w.set || !w.nsch) continue;
            int mask = 0;
            for (int q = 0; q < w.nsch; ++q) mask |= ip.sch[w.sch[q]].v_real << q;
            if (w.nsch == 2) m2 &= mask; else m1 &= mask;
        }
        if (!((m2 == 0 && m1 == 0) || (m2 == 2 && m1 == 0) || (m2 == 3 && m1 == 1))) { m2 = 0; m1 = 0; }   // compiled combinations
    }
    if (c->perf_mode == 1) { m2 = -1; m1 = -1; }                 // factored mode: the main kernel runs EST units only
    void (*main_kernel)(IcParams) =
        m2 < 0 ? (p4s == 4 ? k_ic_main<4, -1, -1> : k_ic_main<8, -1, -1>)
               : p4s == 4 ? (m2 == 2 ? k_ic_main<4, 2, 0> : (m2 == 3 ? k_ic_main<4, 3, 1> : k_ic_main<4, 0, 0>))
                          : (m2 == 2 ? k_ic_main<8, 2, 0> : (m2 == 3 ? k_ic_main<8, 3, 1> : k_ic_main<8, 0, 0>));
    const int cfg_id = p4s * 100 + (m2 + 1) * 10 + (m1 + 1) + (use_post ? 1000 : 0);
    if (c->ic_grid == 0 || c->ic_smem != main_smem + post_smem || c->ic_cfg != cfg_id) {    // persistent main grid: one wave of resident CTAs
        c->ic_cfg = cfg_id;
        CK(cudaFuncSetAttribute(main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CK(cudaFuncSetAttribute(k_ic_light, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CK(cudaFuncSetAttribute(post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        int per_sm = 0, per_sm_light = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, main_kernel, ic_threads, main_smem));
        if (use_post) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_light, post_kernel, POST_THREADS, post_smem));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_light, k_ic_light, IC_LIGHT_THREADS, light_smem));
        if (per_sm < 1 || per_sm_light < 1)
            return fail(CHEST_ERR_STATE, "the IC kernels do not fit on an SM (too many pilots for the shared tables)");
        c->ic_grid = per_sm * c->n_sm; c->ic_light_grid = per_sm_light * c->n_sm; c->ic_smem = main_smem + post_smem;
    }
    CK(c->queue.alloc(32));
    CK(cudaMemsetAsync(c->queue.p, 0, 32 * sizeof(unsigned int), st));
    ip.queue = c->queue.p; ip.n_units = c->n_ctas; ip.n_units_main = c->n_est_units;
    size_t n_err = (size_t)n_rep * S * (n_iter + 1) * 12;
    uint32_t* err = err_dev ? err_dev : c->err.p;
    CK(cudaMemsetAsync(err, 0, n_err * sizeof(uint32_t), st));
    ip.err = err;
    for (int wfi = 0; wfi < 2; ++wfi) { ip.perf_zw[wfi] = c->wf[wfi].det_on ? c->wf[wfi].zw_g.p : nullptr; ip.perf_zw_stride[wfi] = c->wf[wfi].zw_stride; }
    ip.chain_colmajor = (c->chain_cm ? 1 : 0) | (c->chain_cm_perf ? 2 : 0);
    ip.est_fact_mask = 0;
    for (int si = 0; si < 3; ++si) if (c->est_fact[si] && n_iter > 0) ip.est_fact_mask |= 1 << si;
    if (ip.est_fact_mask && use_post) return fail(CHEST_ERR_STATE, "the factored estimator runs with k_ic_light (unset CHEST_LIGHT)");
    ip.mse = nullptr;
    if (c->mse_on) {
        if (use_post) return fail(CHEST_ERR_STATE, "MSE accumulation is implemented in k_ic_light (unset CHEST_LIGHT=post)");
        const size_t n_mse = (size_t)n_rep * S * (n_iter + 1) * 3;
        CK(c->mse.alloc(n_mse));
        CK(cudaMemsetAsync(c->mse.p, 0, n_mse * sizeof(double), st));
        ip.mse = c->mse.p;
    }
    const char* trace_path = getenv("CHEST_IC_TRACE");             // development: per-CTA timestamps of the last main launch
    if (trace_path) { CK(c->trace.alloc((size_t)c->ic_grid * 8)); CK(cudaMemsetAsync(c->trace.p, 0, (size_t)c->ic_grid * 64, st)); }
    const int n_units = std::max(c->n_ctas, 1);
    // test knobs: cap the grids so that a small batch still makes every persistent CTA of k_ic_main pull several
    // units from the queue and every k_ic_light CTA walk its grid-stride loop (tests/test_gpu_parity.py)
    int main_grid = std::min(c->ic_grid, std::max(c->n_est_units, 1));
    int light_grid = std::min(c->ic_light_grid * IC_LIGHT_WAVES, n_units);
    if (const char* e = getenv("CHEST_IC_MAIN_GRID")) main_grid = std::max(1, std::min(main_grid, atoi(e)));
    if (const char* e = getenv("CHEST_IC_LIGHT_GRID")) light_grid = std::max(1, std::min(light_grid, atoi(e)));
    if (c->profiling) CK(cudaEventRecord(c->ev_tw[0], st));
    for (int wfi = 0; wfi < 2; ++wfi) {                            // perfect-CSI twin of whole waveforms: all iterations, one launch
        Waveform& w = c->wf[wfi];
        if (!w.twin_on) continue;
        const ModemDev& md = w.modem;
        PerfTwinParams tp{};
        tp.md = md; tp.nsch = w.nsch; tp.n_snr = S; tp.n_rep = n_rep; tp.n_iter = n_iter; tp.T = c->T; tp.N = N; tp.K = w.K;
        int n_data_max = 0, n_long_max = 0;
        for (int q = 0; q < w.nsch; ++q) {
            tp.sch[q] = ip.sch[w.sch[q]]; tp.scheme_id[q] = w.sch[q];
            n_data_max = std::max(n_data_max, c->sch[w.sch[q]].n_data); n_long_max = std::max(n_long_max, c->sch[w.sch[q]].n_long_rows);
        }
        tp.n_long_max = (n_long_max + 3) & ~3;
        for (int k = 0; k < 2; ++k) tp.cst[k] = c->cst[k].dev;
        tp.htrue = w.htrue.p; tp.h = c->h.p; tp.tap_delay = c->d_tap_delay.p; tp.err = err;
        const size_t smem = ((size_t)2 * md.Ksym * md.nfft + md.nfft) * sizeof(cplx) + (size_t)md.Np * sizeof(double) + (size_t)((md.L + 3) & ~3) * sizeof(int)
                            + (size_t)tp.n_long_max * sizeof(cplx) + (size_t)((w.K + 7) & ~7) * sizeof(unsigned short) + (size_t)((n_data_max + 15) & ~15);
        static size_t attr_dev[64] = {};
        size_t& attr_smem = attr_dev[c->device & 63];
        if (smem > attr_smem) { CK(cudaFuncSetAttribute(k_perfect_twin_fbmc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; }
        k_perfect_twin_fbmc<<<n_rep * w.nsch * S, PERF_FBMC_THREADS, smem, st>>>(tp);
        c->launches++;
        CK(cudaGetLastError());
    }
    for (int it = 0; it <= n_iter; ++it) {
        ip.it = it;
        if (it == 0 && c->perf_mode == 1) { rc = stage_factored_perfect_csi(c, n_rep, 0, n_iter, err, ip); if (rc) return rc; }   // one-tap stage of the detected columns
        if (it == 0 && c->profiling) CK(cudaEventRecord(c->ev_tw[1], st));
        if (it > 0) {                                              // phase B of iteration it
            ip.trace = (trace_path && it == n_iter) ? c->trace.p : nullptr;
            if (c->precision == 1 && c->n_est_units > 0) {
                TcParams tp{};
                tp.it = it; tp.n_iter = n_iter; tp.n_rep = n_rep; tp.K_max = c->K_max; tp.n_items = c->tc_n_items;
                tp.items = c->tc_items.p; tp.cta_ptr = c->tc_cta_ptr.p; tp.ctas = c->ctas.p; tp.scratch = c->scratch.p; tp.status = c->tc_status.p;
                for (int si = 0; si < 3; ++si) {
                    Scheme& s = c->sch[si];
                    if (!s.set) continue;
                    TcScheme& ts = tp.sch[si];
                    ts.K = s.K; ts.P = s.P; ts.n_row_tiles = s.mm[0].tc_row_tiles; ts.y = ip.sch[si].y; ts.hP = ip.sch[si].hP;
                    for (int v = 0; v < 2; ++v) {
                        ts.jlist[v] = s.mm[v].tc_jlist.p; ts.jptr[v] = s.mm[v].tc_jptr.p; ts.a_img[v] = s.mm[v].tc_img.p;
                        ts.a_snr_stride[v] = (long long)s.mm[v].tc_entries * 2 * (c->tc_p8 / 4) * TC_ROWS * 16;
                    }
                }
                if (c->tc_p8 == 16) {
                    static bool attr16_dev[64] = {};
                    bool& attr16 = attr16_dev[c->device & 63];
                    if (!attr16) { CK(cudaFuncSetAttribute(k_ic_est_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGeo<16>::SMEM)); attr16 = true; }
                    k_ic_est_tc<16><<<c->tc_grid, TC_THREADS, TcGeo<16>::SMEM, st>>>(tp);
                } else {
                    static bool attr32_dev[64] = {};
                    bool& attr32 = attr32_dev[c->device & 63];
                    if (!attr32) { CK(cudaFuncSetAttribute(k_ic_est_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGeo<32>::SMEM)); attr32 = true; }
                    k_ic_est_tc<32><<<c->tc_grid, TC_THREADS, TcGeo<32>::SMEM, st>>>(tp);
                }
                c->launches++;
            } else if (c->precision != 1 && c->n_est_units > 0) {
            main_kernel<<<main_grid, ic_threads, main_smem, st>>>(ip);
            c->launches++;
            }
            ip.trace = nullptr;
            if (c->profiling) CK(cudaEventRecord(c->ev_mn[it], st));
            if (ip.est_fact_mask) { rc = stage_factored_estimated_csi(c, n_rep, it, n_iter, ip); if (rc) return rc; }
            if (c->profiling) CK(cudaEventRecord(c->ev_ef[it], st));
            if (c->perf_mode == 1) { rc = stage_factored_perfect_csi(c, n_rep, it, n_iter, err, ip); if (rc) return rc; }
            if (c->profiling) CK(cudaEventRecord(c->ev_ic[2 * it], st));
        }
        // phases C, D, E of iteration it (+ phase A of iteration it+1)
        if (use_post) post_kernel<<<light_grid, POST_THREADS, post_smem, st>>>(ip);
        else k_ic_light<<<light_grid, IC_LIGHT_THREADS, light_smem, st>>>(ip);
        c->launches++;
        CK(cudaGetLastError());
        if (c->profiling) CK(cudaEventRecord(c->ev_ic[2 * it + 1], st));
        if (it == 0 && c->profiling) CK(cudaEventRecord(c->ev[5], st));
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[6], st));
    if (pf_used >= 0) CK(cudaEventRecord(c->pf[pf_used].released, st));        // the set may be refilled after this point
    if (err_host) CK(cudaMemcpyAsync(err_host, err, n_err * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    c->trace_on = trace_path != nullptr; c->pending_iter = n_iter; c->pending_rep = n_rep; c->pending_n_err = n_err;
    if (!wait) {
        if (!err_dev) {                                            // counters travel to a pinned buffer the context owns
            if (c->err_pinned_n < n_err) {
                if (c->err_pinned) cudaFreeHost(c->err_pinned);
                c->err_pinned = nullptr; c->err_pinned_n = 0;
                CK(cudaHostAlloc((void**)&c->err_pinned, n_err * sizeof(uint32_t), cudaHostAllocDefault));
                c->err_pinned_n = n_err;
            }
            CK(cudaMemcpyAsync(c->err_pinned, err, n_err * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        c->pending = true;
        return CHEST_OK;
    }
    return finish_pipeline(c);
}

int finish_pipeline(Ctx* c) {
    cudaStream_t st = c->stream;
    const int n_iter = c->pending_iter, n_rep = c->pending_rep, N = c->N;
    c->pending = false;
    CK(cudaStreamSynchronize(st));
    if (c->precision == 1 && c->tc_status.p) {
        int flag = 0;
        CK(cudaMemcpy(&flag, c->tc_status.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (flag) { cudaMemset(c->tc_status.p, 0, sizeof(int)); return fail(CHEST_ERR_CUDA, "k_ic_est_tc: the tensor-core pipeline gave up on an mbarrier (protocol error)"); }
    }
    if (c->trace_on) {
        const char* trace_path = getenv("CHEST_IC_TRACE");
        std::vector<unsigned long long> th((size_t)c->ic_grid * 8);
        CK(cudaMemcpy(th.data(), c->trace.p, th.size() * 8, cudaMemcpyDeviceToHost));
        if (trace_path) if (FILE* f = fopen(trace_path, "wb")) { fwrite(th.data(), 8, th.size(), f); fclose(f); }
    }
    if (c->profiling) {
        for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]);
        cudaEventElapsedTime(&c->stage_ms[6], c->ev[0], c->ev[6]);
        // per-kernel times: events after every IC launch (main of iteration it: 2 it, light: 2 it + 1; ev[4] precedes)
        c->kernel_ms[2] = c->kernel_ms[3] = c->kernel_ms[4] = 0; c->est_fact_ms = 0;
        for (int it = 0; it <= n_iter; ++it) {
            float t = 0;
            if (it > 0) {
                cudaEventElapsedTime(&t, c->ev_ic[2 * it - 1], c->ev_mn[it]); c->kernel_ms[2] += t;
                cudaEventElapsedTime(&t, c->ev_mn[it], c->ev_ef[it]); c->est_fact_ms += t;
                cudaEventElapsedTime(&t, c->ev_ef[it], c->ev_ic[2 * it]); c->kernel_ms[4] += t;
            }
            cudaEventElapsedTime(&t, it > 0 ? c->ev_ic[2 * it] : c->ev_tw[1], c->ev_ic[2 * it + 1]);
            c->kernel_ms[3] += t;
        }
        { float t = 0; cudaEventElapsedTime(&t, c->ev_tw[0], c->ev_tw[1]); c->kernel_ms[4] += t; }      // fused perfect-CSI twin
        c->kernel_ms[0] = c->kernel_ms[1] = 0;
        c->kernel_ms[5] = c->perf_mode == 1 ? c->stage_ms[2] : 0;              // factored mode: stage 2 is the diag(D) GEMM
        cudaEventElapsedTime(&c->kernel_ms[6], c->ev[1], c->ev_k1[0]);         // k_synth_h
        cudaEventElapsedTime(&c->kernel_ms[7], c->ev_k1[0], c->ev_k1[1]);      // k_tx_symbols
        c->kernel_ms[8] = c->kernel_ms[9] = 0;
        for (int wfi = 0; wfi < 2; ++wfi) {
            float t = 0;
            cudaEventElapsedTime(&t, c->ev_k1[2 + 3 * wfi], c->ev_k1[3 + 3 * wfi]); c->kernel_ms[8] += t;   // s = G x
            cudaEventElapsedTime(&t, c->ev_k1[3 + 3 * wfi], c->ev_k1[4 + 3 * wfi]); c->kernel_ms[9] += t;   // r0 = H s
        }
        if (n_rep > 1 && c->perf_mode != 1)
            for (int wfi = 0; wfi < 2; ++wfi)
                if (c->wf[wfi].set && c->wf[wfi].nsch) {
                    float t = 0;
                    cudaEventElapsedTime(&t, c->ev_hg[2 * wfi + 1], c->ev_gd[wfi]);
                    c->kernel_ms[1] += t;
                }
        c->hg_ms = 0; c->hg_bytes = 0;
        for (int wfi = 0; wfi < 2; ++wfi) {
            Waveform& w = c->wf[wfi];
            if (!w.set || !w.nsch || c->perf_mode == 1 || n_rep <= 1) continue;
            float t = 0;
            cudaEventElapsedTime(&t, c->ev_hg[2 * wfi], c->ev_hg[2 * wfi + 1]);
            c->hg_ms += t;
            // algorithmic bytes of the banded apply to G: write H*G over each column tile's k-range, read h once
            double rows = 0;
            for (size_t tcol = 0; tcol < w.hg_rows.size(); ++tcol) rows += w.hg_rows[tcol];
            c->hg_bytes += 16.0 * (rows * n_rep + (double)c->T * N * n_rep);
        }
    }
    return CHEST_OK;
}

}  // namespace

// ================================================================================ extern "C"
extern "C" {

const char* chest_last_error(void) { return g_err.c_str(); }

int chest_device_info(int device, int* n_sm, int* cc_major, int* cc_minor) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= device) return fail(CHEST_ERR_NO_DEVICE, "no CUDA device");
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, device));
    if (n_sm) *n_sm = pr.multiProcessorCount;
    if (cc_major) *cc_major = pr.major;
    if (cc_minor) *cc_minor = pr.minor;
    return CHEST_OK;
}

int chest_create(int device, uint64_t* handle) {
    ARG(handle);
    int sm = 0, maj = 0, mnr = 0;
    int rc = chest_device_info(device, &sm, &maj, &mnr);
    if (rc) return rc;
    if (maj != 10) return fail(CHEST_ERR_NO_DEVICE, "device is not sm_100 (Blackwell B200); no fallback path exists");
    CK(cudaSetDevice(device));
    Ctx* c = new Ctx();
    if (const char* e = getenv("CHEST_PERF_MODE")) c->perf_mode = (!strcmp(e, "dense") || !strcmp(e, "0")) ? 0 : 1;
    c->device = device; c->n_sm = sm;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& q : c->pf) { CK(cudaEventCreateWithFlags(&q.landed, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&q.released, cudaEventDisableTiming)); }
    for (auto& e : c->ev) CK(cudaEventCreate(&e));
    for (auto& e : c->user_ev) CK(cudaEventCreate(&e));
    for (auto& e : c->ev_hg) CK(cudaEventCreate(&e));
    for (auto& e : c->ev_gd) CK(cudaEventCreate(&e));
    for (auto& e : c->ev_ic) CK(cudaEventCreate(&e));
    for (auto& e : c->ev_mn) CK(cudaEventCreate(&e));
    for (auto& e : c->ev_ef) CK(cudaEventCreate(&e));
    for (auto& e : c->ev_k1) CK(cudaEventCreate(&e));
    for (auto& e : c->ev_tw) CK(cudaEventCreate(&e));
    *handle = (uint64_t)(uintptr_t)c;
    return CHEST_OK;
}

int chest_destroy(uint64_t handle) {
    Ctx* c = from(handle);
    if (!c) return CHEST_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& e : c->ev) cudaEventDestroy(e);
    for (auto& e : c->user_ev) cudaEventDestroy(e);
    for (auto& e : c->ev_hg) cudaEventDestroy(e);
    for (auto& e : c->ev_gd) cudaEventDestroy(e);
    for (auto& e : c->ev_ic) cudaEventDestroy(e);
    for (auto& e : c->ev_mn) cudaEventDestroy(e);
    for (auto& e : c->ev_ef) cudaEventDestroy(e);
    for (auto& e : c->ev_k1) cudaEventDestroy(e);
    for (auto& e : c->ev_tw) cudaEventDestroy(e);
    for (auto& q : c->pf) { cudaEventDestroy(q.landed); cudaEventDestroy(q.released); }
    cudaStreamSynchronize(c->copy_stream);
    if (c->err_pinned) cudaFreeHost(c->err_pinned);
    cudaStream_t s0 = c->stream, s1 = c->copy_stream;
    delete c;                                              // every DevBuf member frees its device memory here
    cudaStreamDestroy(s0);
    cudaStreamDestroy(s1);
    return CHEST_OK;
}

int chest_set_channel(uint64_t handle, int n_samples, int n_taps, const double* pdp, double fd, double dt,
                      int n_paths, int model) {
    Ctx* c = from(handle);
    ARG(c && n_samples > 0 && n_taps > 0 && pdp && n_paths > 0 && dt > 0 && fd >= 0);   // fd = 0: time-invariant, see chest_set_impulse_response
    ARG(model >= CHEST_DOPPLER_JAKES && model <= CHEST_DOPPLER_DISCRETE_UNIFORM);
    if (c->N && c->N != n_samples) return fail(CHEST_ERR_ARG, "Total number of samples must be the same for the channel and every waveform");
    CK(cudaSetDevice(c->device));
    const bool discrete = model >= CHEST_DOPPLER_DISCRETE_JAKES;
    if (discrete && fd > 0 && fd / ((1.0 / dt) / n_samples) <= 0.5) fd = 0;        // FF.m:146-149: velocity too low for a discrete spectrum
    c->N = n_samples; c->Lt = n_taps; c->fD = fd; c->dt = dt; c->paths = n_paths; c->model = model;
    c->n_shift = 0; c->dspec.clear();
    if (discrete && fd > 0) {                                                      // FF.m:151-175
        const double df = (1.0 / dt) / n_samples;
        const int ns = (int)std::ceil(fd / df);
        std::vector<double> pts(2 * ns + 2);
        for (int q = 0; q < 2 * ns + 2; ++q) pts[q] = std::min(fd, std::max(-fd, df * ((q - ns - 1) + 0.5)));
        std::vector<double> sp(2 * ns + 1);
        double tot = 0;
        for (int q = 0; q < 2 * ns + 1; ++q) {
            sp[q] = model == CHEST_DOPPLER_DISCRETE_JAKES ? std::asin(pts[q + 1] / fd) - std::asin(pts[q] / fd) : pts[q + 1] - pts[q];
            tot += sp[q];
        }
        for (double& x : sp) x /= tot;
        c->n_shift = ns; c->dspec = sp;                                            // sp[q]: shift q - ns
    }
    c->tap_delay.clear(); c->tap_amp.clear();
    for (int m = 0; m < n_taps; ++m)
        if (pdp[m] != 0.0) { c->tap_delay.push_back(m); c->tap_amp.push_back(std::sqrt(pdp[m])); }   // FF.m:131,237
    c->T = (int)c->tap_delay.size();
    ARG(c->T > 0);
    CK(c->d_tap_delay.upload(c->tap_delay, c->stream));
    CK(c->d_tap_amp.upload(c->tap_amp, c->stream));
    if (c->n_shift > 0) {      // rows in the reference's order: shifts 0..ns (GaussUncorr1), then -ns..-1 (GaussUncorr2), FF.m:214-217
        const int ns = c->n_shift, nb = 2 * ns + 1;
        std::vector<double> coef((size_t)nb * c->T);
        for (int b = 0; b < nb; ++b) {
            const int shift = b <= ns ? b : b - nb;
            for (int t = 0; t < c->T; ++t) coef[(size_t)b * c->T + t] = std::sqrt(c->dspec[shift + ns]) * c->tap_amp[t] / std::sqrt(2.0);
        }
        CK(c->d_dcoef.upload(coef, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    c->chan_set = true; c->finalized = false;
    // the pseudo-channels / estimator factors belong to the previous channel statistics: the setup of the new channel
    // (chest_setup_correlations + chest_build_mmse, or the factor uploads) brings them back
    for (auto& w : c->wf) w.mq_P = 0;
    for (auto& s : c->sch) for (auto& m : s.mm) m.rinv_set = false;
    c->ctas_for_batch = -1;
    return CHEST_OK;
}

int chest_set_waveform(uint64_t handle, int wfi, int n_samples, int K, const double* G, const double* Q) {
    Ctx* c = from(handle);
    ARG(c && (wfi == 0 || wfi == 1) && n_samples > 0 && K > 0 && G && Q);
    if (c->N && c->N != n_samples) return fail(CHEST_ERR_ARG, "Total number of samples must be the same for the channel and every waveform");   // DS.m:79-81
    CK(cudaSetDevice(c->device));
    Waveform& w = c->wf[wfi];
    const int N = n_samples;
    c->N = N;
    w.K = K;
    CK(w.G.upload(reinterpret_cast<const cplx*>(G), (size_t)N * K, c->stream));
    w.Gh.assign(reinterpret_cast<const cplx*>(G), reinterpret_cast<const cplx*>(G) + (size_t)N * K);
    w.Qh.assign(reinterpret_cast<const cplx*>(Q), reinterpret_cast<const cplx*>(Q) + (size_t)N * K);
    CK(w.Q.upload(reinterpret_cast<const cplx*>(Q), (size_t)N * K, c->stream));
    {   // operand planes of Q^H for the three-multiplication GEMMs (K2, K3a): conj(q) = (re, -im)
        const int Np = (N + 1) & ~1;
        const cplx* q = reinterpret_cast<const cplx*>(Q);
        std::vector<cplx> q1((size_t)Np * K, cmake(0.0, 0.0));
        std::vector<double> q2((size_t)Np * K, 0.0);
        for (int i = 0; i < K; ++i)
            for (int n = 0; n < N; ++n) {
                const cplx v = q[(size_t)n + (size_t)N * i];
                q1[(size_t)n + (size_t)Np * i] = cmake(v.x, v.x - v.y);
                q2[(size_t)n + (size_t)Np * i] = v.y;
            }
        CK(w.Q1.upload(q1, c->stream)); CK(w.Q2.upload(q2, c->stream));
    }
    std::vector<cplx> gt((size_t)N * K);
    const cplx* g = reinterpret_cast<const cplx*>(G);
    for (int j = 0; j < K; ++j)
        for (int n = 0; n < N; ++n) gt[(size_t)j + (size_t)K * n] = g[(size_t)n + (size_t)N * j];
    CK(w.Gt.upload(gt, c->stream));
    {   // planes of G for the ring GEMM s = G v of the factored perfect-CSI pass
        const int Kp = (K + 1) & ~1;
        std::vector<cplx> g1((size_t)Kp * N, cmake(0.0, 0.0));
        std::vector<double> g2((size_t)Kp * N, 0.0);
        for (int j = 0; j < K; ++j)
            for (int n = 0; n < N; ++n) {
                const cplx v = g[(size_t)n + (size_t)N * j];
                g1[(size_t)j + (size_t)Kp * n] = cmake(v.x, v.x + v.y);
                g2[(size_t)j + (size_t)Kp * n] = -v.y;
            }
        CK(w.Gt1.upload(g1, c->stream)); CK(w.Gt2.upload(g2, c->stream));
    }
    support_ranges(G, N, K, w.g_lo, w.g_hi);
    support_ranges(Q, N, K, w.q_lo, w.q_hi);
    std::vector<int> lo, hi;
    {   // tile size: fewest padded flops of D = Q^H (H G) (k-ranges are unions over the tile's columns)
        double best = -1;
        for (int t : {64, 48}) {
            std::vector<int> ql, qh, gl, gh;
            tile_ranges(w.q_lo, w.q_hi, t, 0, N, ql, qh);
            tile_ranges(w.g_lo, w.g_hi, t, 8, N, gl, gh);
            double f = 0;
            for (size_t a = 0; a < ql.size(); ++a)
                for (size_t b = 0; b < gl.size(); ++b) {
                    int l = std::max(ql[a], gl[b]), h = std::min(qh[a], gh[b]);
                    if (h > l) f += (double)t * t * (((h - l) + 15) / 16) * 16;
                }
            if (best < 0 || f < best) { best = f; w.tile = t; }
        }
        if (const char* e = getenv("CHEST_GEMM_TILE")) w.tile = atoi(e) == 48 ? 48 : 64;
    }
    tile_ranges(w.q_lo, w.q_hi, w.tile, 0, N, lo, hi);
    CK(w.q_klo.upload(lo, c->stream)); CK(w.q_khi.upload(hi, c->stream));
    tile_ranges(w.q_lo, w.q_hi, 8, 0, N, lo, hi);
    CK(w.q8_klo.upload(lo, c->stream)); CK(w.q8_khi.upload(hi, c->stream));
    CK(w.q_lo_d.upload(w.q_lo, c->stream)); CK(w.q_hi_d.upload(w.q_hi, c->stream));
    // row supports of G (which symbols j touch sample n) for s = G x
    std::vector<int> rlo(N, K), rhi(N, 0);
    for (int j = 0; j < K; ++j)
        for (int n = w.g_lo[j]; n < w.g_hi[j]; ++n) { rlo[n] = std::min(rlo[n], j); rhi[n] = std::max(rhi[n], j + 1); }
    tile_ranges(rlo, rhi, w.tile, 0, K, lo, hi);
    CK(w.gt_klo.upload(lo, c->stream)); CK(w.gt_khi.upload(hi, c->stream));
    tile_ranges(rlo, rhi, 8, 0, K, lo, hi);
    CK(w.gt8_klo.upload(lo, c->stream)); CK(w.gt8_khi.upload(hi, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    w.set = true; c->finalized = false; w.pf_state = 0; c->ctas_for_batch = -1;
    return CHEST_OK;
}

int chest_set_constellation(uint64_t handle, int which, int order, const double* sym, const uint8_t* bitmap) {
    Ctx* c = from(handle);
    ARG(c && (which == 0 || which == 1) && order >= 2 && sym && bitmap);
    // decided words travel as one byte per data symbol inside the IC kernels (k_ic_light, zw[])
    if (order > 256) return fail(CHEST_ERR_ARG, "constellation order > 256 is not supported (decided words are stored as bytes)");
    CK(cudaSetDevice(c->device));
    Constellation& k = c->cst[which];
    int nb = 0;
    while ((1 << nb) < order) ++nb;
    ARG((1 << nb) == order);
    k.order = order; k.nbits = nb; k.is_qam = (which == CHEST_CONST_QAM);
    // SymbolMapping is sorted by bit value, so the row index must equal bi2de(BitMapping(row,:)) (SC.m:64-66)
    for (int m = 0; m < order; ++m) {
        int w = 0;
        for (int t = 0; t < nb; ++t) w |= (bitmap[m + (size_t)order * t] & 1) << t;
        if (w != m) return fail(CHEST_ERR_ARG, "BitMapping rows are not sorted by bit value");
    }
    std::vector<cplx> s(order), pil(order);
    for (int m = 0; m < order; ++m) {
        s[m] = cmake(sym[2 * m], sym[2 * m + 1]);
        double a = std::hypot(s[m].x, s[m].y);
        pil[m] = cmake(s[m].x / a, s[m].y / a);                       // xP./abs(xP), DS.m:366,368
    }
    k.real = true;
    for (int m = 0; m < order; ++m) if (s[m].y != 0.0 || pil[m].y != 0.0) k.real = false;
    std::vector<double> lev;
    for (int m = 0; m < order; ++m) lev.push_back(s[m].x);
    std::sort(lev.begin(), lev.end());
    lev.erase(std::unique(lev.begin(), lev.end()), lev.end());
    k.n_axis = (int)lev.size();
    ARG(k.n_axis >= 2);
    std::vector<int> grid;
    if (!k.is_qam) {
        ARG(k.n_axis == order);
        grid.assign(order, -1);
        for (int m = 0; m < order; ++m) {
            ARG(s[m].y == 0.0);
            int t = (int)(std::lower_bound(lev.begin(), lev.end(), s[m].x) - lev.begin());
            grid[t] = m;
        }
    } else {
        ARG(k.n_axis * k.n_axis == order);
        grid.assign(order, -1);
        for (int m = 0; m < order; ++m) {
            int ti = (int)(std::lower_bound(lev.begin(), lev.end(), s[m].x) - lev.begin());
            auto itq = std::lower_bound(lev.begin(), lev.end(), s[m].y);
            ARG(itq != lev.end() && *itq == s[m].y);
            grid[ti * k.n_axis + (int)(itq - lev.begin())] = m;
        }
    }
    for (int v : grid) ARG(v >= 0);
    CK(k.symbol.upload(s, c->stream)); CK(k.pilot.upload(pil, c->stream));
    CK(k.level.upload(lev, c->stream)); CK(k.word_of_grid.upload(grid, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    k.dev.order = order; k.dev.nbits = nb; k.dev.n_axis = k.n_axis; k.dev.is_qam = k.is_qam;
    k.dev.inv_step = (double)(k.n_axis - 1) / (lev.back() - lev.front());
    k.dev.symbol = k.symbol.p; k.dev.pilot = k.pilot.p; k.dev.level = k.level.p; k.dev.word_of_grid = k.word_of_grid.p;
    k.set = true; c->finalized = false;
    return CHEST_OK;
}

int chest_set_scheme(uint64_t handle, int si, int wfi, int k_in, int P, int n_data, const int64_t* jc,
                     const int32_t* ir, const double* val, const int32_t* pilot_pos, const int32_t* data_pos,
                     double kappa, double dpr, int detect, int constellation, const uint8_t* considered) {
    Ctx* c = from(handle);
    ARG(c && si >= 0 && si < 3 && (wfi == 0 || wfi == 1) && jc && ir && val && pilot_pos && considered);
    ARG(constellation == CHEST_CONST_PAM || constellation == CHEST_CONST_QAM);
    ARG((c->wf[wfi].set || c->wf[wfi].modem_set) && c->cst[constellation].set);
    ARG(detect >= 0 && detect <= 2 && (detect == CHEST_DETECT_DESPREAD_REAL || data_pos));
    ARG(P > 0 && P <= 128 && n_data > 0 && k_in >= P + n_data && kappa > 0 && dpr > 0);
    CK(cudaSetDevice(c->device));
    Scheme& s = c->sch[si];
    const int K = c->wf[wfi].K;
    s.waveform = wfi; s.K = K; s.K_in = k_in; s.P = P; s.n_data = n_data; s.detect = detect;
    s.constellation = constellation; s.kappa = kappa; s.dpr = dpr;
    const int nb = c->cst[constellation].nbits;
    s.n_bits = n_data * nb;
    ARG(nb <= 32);
    const int64_t nnz = jc[k_in];
    s.c_nnz = nnz;
    const cplx* v = reinterpret_cast<const cplx*>(val);
    s.c_real = true;                                       // exactly real precoder: v = C z is real for a real constellation
    for (int64_t e = 0; e < nnz; ++e) if (v[e].y != 0.0) { s.c_real = false; break; }
    // CSC (columns of C) for the de-spreading C' x, CSR (rows of C) for the precoding C z
    std::vector<int> colptr(k_in + 1), rows(nnz), rowptr(K + 1, 0), cols(nnz);
    std::vector<cplx> vcsc(nnz), vcsr(nnz);
    for (int k = 0; k <= k_in; ++k) colptr[k] = (int)jc[k];
    for (int64_t e = 0; e < nnz; ++e) { ARG(ir[e] >= 0 && ir[e] < K); rows[e] = ir[e]; vcsc[e] = v[e]; rowptr[ir[e] + 1]++; }
    for (int i = 0; i < K; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int> fill(rowptr.begin(), rowptr.end() - 1);
    for (int k = 0; k < k_in; ++k)
        for (int e = colptr[k]; e < colptr[k + 1]; ++e) { int d = fill[rows[e]]++; cols[d] = k; vcsr[d] = vcsc[e]; }
    CK(s.c_rowptr.upload(rowptr, c->stream)); CK(s.c_col.upload(cols, c->stream)); CK(s.c_val.upload(vcsr, c->stream));
    {   // ELL-1 view: rows with at most one entry are applied element-wise, the others are "long rows"
        std::vector<int> col0(K, -1), longr;
        std::vector<cplx> val0(K, cmake(0.0, 0.0));
        for (int i = 0; i < K; ++i) {
            int n = rowptr[i + 1] - rowptr[i];
            if (n == 1) { col0[i] = cols[rowptr[i]]; val0[i] = vcsr[rowptr[i]]; }
            else if (n > 1) { col0[i] = -2; longr.push_back(i); }
        }
        s.n_long_rows = (int)longr.size();
        {   // long rows as DMMA tiles: 8 rows per tile, the union of their columns packed four per k-step
            const int nt = ((int)longr.size() + 7) / 8;
            std::vector<int> lptr(1, 0), kcol;
            std::vector<cplx> frag;
            for (int t = 0; t < nt; ++t) {
                std::vector<int> u;
                for (int g = 0; g < 8 && t * 8 + g < (int)longr.size(); ++g) {
                    int i = longr[t * 8 + g];
                    u.insert(u.end(), cols.begin() + rowptr[i], cols.begin() + rowptr[i + 1]);
                }
                std::sort(u.begin(), u.end());
                u.erase(std::unique(u.begin(), u.end()), u.end());
                const int n_real = (int)u.size();
                while (u.size() % 4) u.push_back(u[0]);               // padding columns carry zero coefficients
                for (size_t st = 0; st < u.size() / 4; ++st) {
                    for (int q = 0; q < 4; ++q) kcol.push_back(u[st * 4 + q]);
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, q = lane & 3;
                        cplx a = cmake(0.0, 0.0);
                        if (t * 8 + g < (int)longr.size() && (int)(st * 4 + q) < n_real) {
                            const int i = longr[t * 8 + g], col = u[st * 4 + q];
                            for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) if (cols[e] == col) { a.x += vcsr[e].x; a.y += vcsr[e].y; }
                        }
                        frag.push_back(a);
                    }
                }
                lptr.push_back((int)kcol.size() / 4);
            }
            s.n_lr_tiles = nt;
            if (kcol.empty()) { kcol.assign(4, 0); frag.assign(32, cmake(0.0, 0.0)); }
            CK(s.lr_ptr.upload(lptr, c->stream)); CK(s.lr_kcol.upload(kcol, c->stream)); CK(s.lr_frag.upload(frag, c->stream));
        }
        if (longr.empty()) longr.push_back(0);
        CK(s.row_col0.upload(col0, c->stream)); CK(s.row_val0.upload(val0, c->stream)); CK(s.long_rows.upload(longr, c->stream));
    }
    CK(s.ct_colptr.upload(colptr, c->stream)); CK(s.ct_row.upload(rows, c->stream)); CK(s.ct_val.upload(vcsc, c->stream));
    std::vector<int> pp(pilot_pos, pilot_pos + P);
    for (int x : pp) ARG(x >= 0 && x < K);
    CK(s.pilot_pos.upload(pp, c->stream));
    {
        std::vector<int> p2d(K, -1);                       // position -> data symbol read off it (select schemes)
        if (data_pos) {
            std::vector<int> dp(data_pos, data_pos + n_data);
            for (int d = 0; d < n_data; ++d) {
                ARG(dp[d] >= 0 && dp[d] < K);
                if (p2d[dp[d]] >= 0) return fail(CHEST_ERR_ARG, "two data symbols are read off the same position");
                p2d[dp[d]] = d;
            }
            CK(s.data_pos.upload(dp, c->stream));
        }
        CK(s.pos2data.upload(p2d, c->stream));
    }
    std::vector<uint32_t> mask(n_data, 0);
    s.considered.assign(considered, considered + (size_t)n_data * nb);
    s.n_bits_edge = 0;
    for (int d = 0; d < n_data; ++d)
        for (int t = 0; t < nb; ++t)
            if (considered[(size_t)d * nb + t]) { mask[d] |= 1u << t; s.n_bits_edge++; }
    CK(s.edge_mask.upload(mask, c->stream));
    {   // row descriptors of k_ic_post
        std::vector<int> col0(K, -1), d_direct(K, -1);
        std::vector<char> need(K, 0);
        for (int i = 0; i < K; ++i) {
            const int n = rowptr[i + 1] - rowptr[i];
            if (n == 1) col0[i] = cols[rowptr[i]]; else if (n > 1) col0[i] = -2;
        }
        std::vector<int> multi;
        if (detect == CHEST_DETECT_DESPREAD_REAL) {
            // a data symbol is decided at its own row when its spreading set is that single row and the row holds nothing else
            for (int d = 0; d < n_data; ++d) {
                const int k = P + d, e0 = colptr[k], e1 = colptr[k + 1];
                if (e1 - e0 == 1 && col0[rows[e0]] == k) d_direct[rows[e0]] = d;
                else { multi.push_back(d); for (int e = e0; e < e1; ++e) need[rows[e]] = 1; }
            }
        } else {
            std::vector<int> dp(data_pos, data_pos + n_data);
            for (int d = 0; d < n_data; ++d) d_direct[dp[d]] = d;
        }
        bool fuse = true;
        for (int i = 0; i < K; ++i) {
            if (col0[i] >= P && d_direct[i] != col0[i] - P) fuse = false;     // the row's entry refers to a symbol decided elsewhere
            if (d_direct[i] >= 0 && col0[i] >= 0 && col0[i] < P) fuse = false; // a decided row whose entry refers to a pilot
            if (need[i] && col0[i] != -2) fuse = false;                        // x-hat and v of the row would share a slot
        }
        if (getenv("CHEST_POST_NOFUSE")) fuse = false;                         // test knob: exercise the separate phase-A pass
        std::vector<int4> ri(K);
        for (int i = 0; i < K; ++i) ri[i] = make_int4(d_direct[i], col0[i], d_direct[i] >= 0 ? (int)mask[d_direct[i]] : 0, need[i]);
        s.n_multi = (int)multi.size(); s.fuse_ok = fuse;
        if (multi.empty()) multi.push_back(0);
        CK(s.rowinfo.upload(ri, c->stream)); CK(s.multi_d.upload(multi, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    s.set = true; c->finalized = false;
    return CHEST_OK;
}

int chest_set_snr(uint64_t handle, int n_snr, const double* pn) {
    Ctx* c = from(handle);
    ARG(c && n_snr > 0 && pn);
    CK(cudaSetDevice(c->device));
    c->S = n_snr; c->pn.assign(pn, pn + n_snr);
    std::vector<double> sc(n_snr);
    for (int i = 0; i < n_snr; ++i) sc[i] = std::sqrt(pn[i] / 2.0);      // DS.m:399
    CK(c->d_noise_scale.upload(sc, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->finalized = false;
    return CHEST_OK;
}

int chest_set_mmse(uint64_t handle, int si, int variant, int n_snr, const int64_t* jc, const int64_t* ir,
                   const double* val) {
    Ctx* c = from(handle);
    ARG(c && si >= 0 && si < 3 && (variant == 0 || variant == 1) && jc && ir && val);
    ARG(c->sch[si].set && n_snr == c->S);
    CK(cudaSetDevice(c->device));
    Scheme& s = c->sch[si];
    MmseVariant& m = s.mm[variant];
    m.rinv_set = false; c->ctas_for_batch = -1;               // an uploaded W carries no factors (chest_set_estimator_factors adds them)
    const int K = s.K, P = s.P, P4 = (P + 3) / 4, RT = (K + 7) / 8, ND = 2 * K - 1;
    const int64_t K2 = (int64_t)K * K;
    const cplx* v = reinterpret_cast<const cplx*>(val);
    // pass 1: which (row tile, diagonal offset) pairs hold any off-diagonal non-zero
    std::vector<int> lut((size_t)RT * ND, -1);
    for (int64_t e = 0; e < jc[n_snr]; ++e) {
        int64_t r = ir[e];
        ARG(r >= 0 && r < K2 * P);
        int64_t rem = r % K2;
        int j = (int)(rem / K), i = (int)(rem % K);
        if (i != j) lut[(size_t)(i >> 3) * ND + (j - i + K - 1)] = 0;
    }
    std::vector<int> tptr(RT + 1, 0), tdel;
    for (int rt = 0; rt < RT; ++rt) {
        for (int d = 0; d < ND; ++d)
            if (lut[(size_t)rt * ND + d] == 0) { lut[(size_t)rt * ND + d] = (int)tdel.size(); tdel.push_back(d - (K - 1)); }
        tptr[rt + 1] = (int)tdel.size();
    }
    m.n_tiles = (int)tdel.size();
    m.tc_packed = false;                                       // the split-BF16 operand images follow W
    CK(m.tile_ptr.upload(tptr, c->stream));
    CK(m.tile_delta.upload(tdel.empty() ? std::vector<int>(1, 0) : tdel, c->stream));
    m.frag.clear(); m.diag.clear();
    m.frag.resize(n_snr); m.diag.resize(n_snr);
    std::vector<WTiles> table(n_snr);
    std::vector<cplx> frag((size_t)std::max(m.n_tiles, 1) * P4 * 32), dg((size_t)K * P);
    std::vector<cplx> dfrag((size_t)n_snr * RT * P4 * 32, cmake(0.0, 0.0));
    std::vector<char> pair_seen;
    m.nnz_offdiag_pairs = 0;
    for (int snr = 0; snr < n_snr; ++snr) {
        std::fill(frag.begin(), frag.end(), cmake(0.0, 0.0));
        std::fill(dg.begin(), dg.end(), cmake(0.0, 0.0));
        for (int64_t e = jc[snr]; e < jc[snr + 1]; ++e) {
            int64_t r = ir[e];
            int p = (int)(r / K2);
            int64_t rem = r % K2;
            int j = (int)(rem / K), i = (int)(rem % K);
            if (i == j) {
                dg[(size_t)i * P + p] = v[e];
                dfrag[(((size_t)snr * RT + (i >> 3)) * P4 + (p >> 2)) * 32 + (i & 7) * 4 + (p & 3)] = v[e];
                continue;
            }
            int t = lut[(size_t)(i >> 3) * ND + (j - i + K - 1)];
            frag[((size_t)t * P4 + (p >> 2)) * 32 + (i & 7) * 4 + (p & 3)] = v[e];
        }
        CK(m.frag[snr].upload(frag, c->stream));
        CK(m.diag[snr].upload(dg, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        table[snr].frag = m.frag[snr].p; table[snr].diag = m.diag[snr].p;
    }
    // exact count of off-diagonal (i,j) pairs with any non-zero pilot weight (work model)
    {
        std::vector<char> seen((size_t)K2, 0);
        for (int64_t e = 0; e < jc[n_snr]; ++e) seen[(size_t)(ir[e] % K2)] = 1;
        int64_t cnt = 0;
        for (int64_t q = 0; q < K2; ++q) if (seen[q] && (q / K) != (q % K)) ++cnt;
        m.nnz_offdiag_pairs = cnt;
    }
    CK(m.table.upload(table, c->stream));
    CK(m.diag_frag.upload(dfrag, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    m.set = true; c->finalized = false;
    return CHEST_OK;
}

// ---------------------------------------------------------------- setup on the device (DS.m:208-313)
int chest_setup_correlations(uint64_t handle, int wfi, int n_pilots, const int32_t* pilot_pos, const double* time_corr,
                             double zero_threshold, double* R_hP_out, int64_t* n_support) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG((wfi == 0 || wfi == 1) && c->wf[wfi].set && n_pilots >= 1 && n_pilots <= c->max_batch && pilot_pos && time_corr && R_hP_out);
    CK(cudaSetDevice(c->device));
    Waveform& w = c->wf[wfi];
    const int N = c->N, K = w.K, T = c->T, P = n_pilots, max_delay = std::max(1, c->tap_delay.back());
    const size_t RT8 = ((size_t)(K + 7) / 8) * 8, n_e = RT8 * K;
    std::vector<int> pil(pilot_pos, pilot_pos + P);
    for (int x : pil) ARG(x >= 0 && x < K);
    std::vector<double> tp(T);
    for (int t = 0; t < T; ++t) tp[t] = c->tap_amp[t] * c->tap_amp[t];
    // scratch of this call, kept in the context: a velocity sweep calls it once per velocity
    DevBuf<int>& d_pil = c->setup_pil; DevBuf<double>& d_rt = c->setup_rt; DevBuf<double>& d_tp = c->setup_tp;
    DevBuf<cplx>& corner = c->setup_corner; DevBuf<cplx>& rhp = c->setup_rhp;
    cudaStream_t st = c->stream;
    CK(d_pil.upload(pil, st)); CK(d_rt.upload(time_corr, (size_t)2 * N - 1, st)); CK(d_tp.upload(tp, st));
    CK(w.g_lo_d.upload(w.g_lo, st)); CK(w.g_hi_d.upload(w.g_hi, st));
    CK(corner.alloc((size_t)P * T * max_delay)); CK(rhp.alloc((size_t)P * P));
    CK(cudaMemsetAsync(corner.p, 0, sizeof(cplx) * P * T * max_delay, st));
    dim3 g1((N + 127) / 128, T, P);
    k_pseudo_channel<<<g1, 128, 0, st>>>(c->h.p, corner.p, w.G.p, w.Q.p, d_pil.p, w.g_lo_d.p, w.g_hi_d.p, d_rt.p,
                                         c->d_tap_delay.p, d_tp.p, N, T, max_delay);
    c->launches++;
    CK(cudaGetLastError());
    {   // the pseudo-channels are the factors of the estimated channel H-hat = sum_q g_q M_q (factored estimator): keep them
        CK(w.Mq.alloc((size_t)P * T * N));
        CK(cudaMemcpyAsync(w.Mq.p, c->h.p, sizeof(cplx) * P * T * N, cudaMemcpyDeviceToDevice, st));
        std::vector<cplx> ch((size_t)P * T * max_delay);
        CK(cudaMemcpyAsync(ch.data(), corner.p, sizeof(cplx) * ch.size(), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        w.mq_corner = false;
        for (const cplx& x : ch) if (x.x != 0.0 || x.y != 0.0) { w.mq_corner = true; break; }
        w.mq_P = P;
        CK(c->zmax.alloc(1)); CK(cudaMemsetAsync(c->zmax.p, 0, sizeof(unsigned long long), st));
    }
    c->cur_batch = P;
    rc = stage_transmission_matrix(c, wfi, P, 0);                   // D_p = Q^H M_p G for every pilot: K1 + K2
    if (rc) return rc;
    CK(w.Rsup.alloc((size_t)P * n_e));
    dim3 g2((unsigned)((n_e + 255) / 256), P);
    k_rsup_finish<<<g2, 256, 0, st>>>(w.Rsup.p, w.D.p, corner.p, w.G.p, w.Q.p, c->d_tap_delay.p, N, K, T, max_delay, 0.0, nullptr);
    k_rhp_gather<<<P, ((P + 31) / 32) * 32, 0, st>>>(rhp.p, w.Rsup.p, d_pil.p, K, P);
    k_rsup_finish<<<g2, 256, 0, st>>>(w.Rsup.p, w.Rsup.p, corner.p, w.G.p, w.Q.p, c->d_tap_delay.p, N, K, 0, max_delay, zero_threshold, c->zmax.p);
    c->launches += 3;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(R_hP_out, rhp.p, sizeof(cplx) * P * P, cudaMemcpyDeviceToHost, st));
    { unsigned long long zb = 0; CK(cudaMemcpyAsync(&zb, c->zmax.p, sizeof(zb), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
      long long zl = (long long)zb; double zd; memcpy(&zd, &zl, sizeof(zd)); w.rsup_zeroed_max = zd; }
    CK(cudaStreamSynchronize(st));
    w.rsup_P = P; w.rsup_thr = zero_threshold; c->ctas_for_batch = -1;
    if (n_support) {                                                // entries of the K x K grid with any non-zero pilot weight
        DevBuf<int>& mask = c->setup_mask; DevBuf<cplx>& eye = c->setup_eye;       // kept in the context: no cudaMalloc / cudaFree per velocity
        CK(mask.alloc(n_e)); CK(cudaMemsetAsync(mask.p, 0, n_e * sizeof(int), st));
        std::vector<cplx> id((size_t)P * P, cmake(0.0, 0.0));
        for (int p = 0; p < P; ++p) id[(size_t)p * P + p] = cmake(1.0, 0.0);
        CK(eye.upload(id, st));
        dim3 g3((unsigned)((n_e * P + 255) / 256), 1);
        k_w_mask<<<g3, 256, 0, st>>>(mask.p, w.Rsup.p, eye.p, K, P, 1e-300, nullptr);
        c->launches++;
        std::vector<int> hm(n_e);
        CK(cudaMemcpyAsync(hm.data(), mask.p, n_e * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int64_t cnt = 0;
        for (int v : hm) cnt += v;
        *n_support = cnt;
    }
    return CHEST_OK;
}

int chest_build_mmse(uint64_t handle, int si, int variant, int n_snr, const double* R_inv, double zero_threshold) {
    Ctx* c = from(handle);
    ARG(c && si >= 0 && si < 3 && (variant == 0 || variant == 1) && R_inv);
    ARG(c->sch[si].set && n_snr == c->S);
    CK(cudaSetDevice(c->device));
    Scheme& s = c->sch[si];
    Waveform& w = c->wf[s.waveform];
    if (!w.Rsup.p || w.rsup_P != s.P) return fail(CHEST_ERR_STATE, "chest_build_mmse: call chest_setup_correlations for this waveform first");
    MmseVariant& m = s.mm[variant];
    const int K = s.K, P = s.P, P4 = (P + 3) / 4, RT = (K + 7) / 8, ND = 2 * K - 1;
    const size_t n_e = (size_t)RT * 8 * K;
    cudaStream_t st = c->stream;
    DevBuf<cplx>& rinv = c->setup_rinv; DevBuf<int>& mask = c->setup_mask;
    CK(rinv.upload(reinterpret_cast<const cplx*>(R_inv), (size_t)n_snr * P * P, st));
    CK(mask.alloc(n_e)); CK(cudaMemsetAsync(mask.p, 0, n_e * sizeof(int), st));
    dim3 g1((unsigned)((n_e * P + 255) / 256), n_snr);
    CK(m.rinv.upload(reinterpret_cast<const cplx*>(R_inv), (size_t)n_snr * P * P, st));      // kept: g = Rinv hP of the factored estimator
    CK(c->zmax.alloc(1)); CK(cudaMemsetAsync(c->zmax.p, 0, sizeof(unsigned long long), st));
    if (P <= 16) { dim3 gm((unsigned)((n_e + 255) / 256), n_snr); k_w_mask_rows<16><<<gm, 256, (size_t)P * P * sizeof(cplx), st>>>(mask.p, w.Rsup.p, rinv.p, K, P, zero_threshold, c->zmax.p); }
    else if (P <= 32) { dim3 gm((unsigned)((n_e + 127) / 128), n_snr); k_w_mask_rows<32><<<gm, 128, (size_t)P * P * sizeof(cplx), st>>>(mask.p, w.Rsup.p, rinv.p, K, P, zero_threshold, c->zmax.p); }
    else k_w_mask<<<g1, 256, 0, st>>>(mask.p, w.Rsup.p, rinv.p, K, P, zero_threshold, c->zmax.p);
    c->launches++;
    CK(cudaGetLastError());
    std::vector<int> hm(n_e);
    unsigned long long zb = 0;
    CK(cudaMemcpyAsync(hm.data(), mask.p, n_e * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&zb, c->zmax.p, sizeof(zb), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    { long long zl = (long long)zb; double zd; memcpy(&zd, &zl, sizeof(zd)); m.w_zeroed_max = zd; m.rinv_set = true; }
    // tile list: (row tile, diagonal offset) pairs with any surviving off-diagonal entry, offsets ascending per row tile
    std::vector<char> lut((size_t)RT * ND, 0);
    int64_t pairs = 0;
    for (int rt = 0; rt < RT; ++rt)
        for (int j = 0; j < K; ++j)
            for (int r = 0; r < 8; ++r) {
                const int i = rt * 8 + r;
                if (i >= K || i == j || !hm[((size_t)rt * K + j) * 8 + r]) continue;
                lut[(size_t)rt * ND + (j - i + K - 1)] = 1; ++pairs;
            }
    std::vector<int> tptr(RT + 1, 0), tdel, trt;
    for (int rt = 0; rt < RT; ++rt) {
        for (int d = 0; d < ND; ++d) if (lut[(size_t)rt * ND + d]) { tdel.push_back(d - (K - 1)); trt.push_back(rt); }
        tptr[rt + 1] = (int)tdel.size();
    }
    m.n_tiles = (int)tdel.size(); m.nnz_offdiag_pairs = pairs;
    if (tdel.empty()) { tdel.push_back(0); trt.push_back(0); }
    DevBuf<int>& d_trt = c->setup_trt;
    CK(m.tile_ptr.upload(tptr, st)); CK(m.tile_delta.upload(tdel, st)); CK(d_trt.upload(trt, st));
    // buffers are kept across rebuilds (cudaMalloc / cudaFree of ~100 buffers cost more than the kernels of a rebuild);
    // a little headroom absorbs the small changes of the tile count from one velocity to the next
    if ((int)m.frag.size() != n_snr) { m.frag.clear(); m.diag.clear(); m.frag.resize(n_snr); m.diag.resize(n_snr); }
    m.tc_packed = false;                                       // the split-BF16 operand images follow W
    std::vector<WTiles> table(n_snr);
    const size_t n_frag = (size_t)std::max(m.n_tiles, 1) * P4 * 32, n_dfrag = (size_t)RT * P4 * 32;
    CK(m.diag_frag.alloc((size_t)n_snr * n_dfrag));
    CK(cudaMemsetAsync(m.diag_frag.p, 0, sizeof(cplx) * n_snr * n_dfrag, st));
    const int64_t n_thr = ((int64_t)m.n_tiles + RT) * 8 * P;
    for (int snr = 0; snr < n_snr; ++snr) {
        if (m.frag[snr].n < n_frag) CK(m.frag[snr].alloc(n_frag + n_frag / 2));      // (a velocity sweep grows the support step by step)
        CK(m.diag[snr].alloc((size_t)K * P));
        CK(cudaMemsetAsync(m.frag[snr].p, 0, sizeof(cplx) * n_frag, st));
        CK(cudaMemsetAsync(m.diag[snr].p, 0, sizeof(cplx) * K * P, st));
        k_w_fill<<<(unsigned)((n_thr + 255) / 256), 256, 0, st>>>(m.frag[snr].p, m.diag[snr].p, m.diag_frag.p + (size_t)snr * n_dfrag,
                                                                  w.Rsup.p, rinv.p + (size_t)snr * P * P, d_trt.p, m.tile_delta.p,
                                                                  m.n_tiles, K, P, P4, zero_threshold);
        c->launches++;
        table[snr].frag = m.frag[snr].p; table[snr].diag = m.diag[snr].p;
    }
    CK(cudaGetLastError());
    CK(m.table.upload(table, st));
    CK(cudaStreamSynchronize(st));
    m.set = true; c->finalized = false; c->ctas_for_batch = -1;
    return CHEST_OK;
}

int chest_release_setup(uint64_t handle) {
    Ctx* c = from(handle);
    ARG(c);
    CK(cudaSetDevice(c->device));
    for (auto& w : c->wf) { w.Rsup.release(); w.rsup_P = 0; }
    return CHEST_OK;
}

int chest_finalize(uint64_t handle, int max_batch) {
    Ctx* c = from(handle);
    ARG(c && max_batch >= 1);
    const bool modem_only = !c->chan_set && (c->wf[0].modem_set || c->wf[1].modem_set) && !c->wf[0].set && !c->wf[1].set;
    if (!c->chan_set && !modem_only) return fail(CHEST_ERR_STATE, "channel not set");
    CK(cudaSetDevice(c->device));
    const int B = max_batch, S = c->S, N = c->N;
    c->K_max = 0;
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.set) continue;
        const int K = w.K, max_delay = c->tap_delay.back();
        std::vector<int> lo, hi;
        tile_ranges(w.g_lo, w.g_hi, w.tile, max_delay, N, lo, hi);
        CK(w.hg_klo.upload(lo, c->stream)); CK(w.hg_khi.upload(hi, c->stream));
        w.hg_rows.assign(K, 0.0);
        for (int j = 0; j < K; ++j) w.hg_rows[j] = std::max(0, hi[j / w.tile] - lo[j / w.tile]);
        tile_ranges(w.g_lo, w.g_hi, 8, max_delay, N, lo, hi);
        CK(w.hg8_klo.upload(lo, c->stream)); CK(w.hg8_khi.upload(hi, c->stream));
        w.pd_built = false;                                    // diag(D) operand: built by the first factored run (ensure_pd)
        {   // work list of K2: tile pairs of D whose Q / H*G supports overlap (the rest of D is structurally zero)
            std::vector<int> ql, qh, gl, gh;
            tile_ranges(w.q_lo, w.q_hi, w.tile, 0, N, ql, qh);
            tile_ranges(w.g_lo, w.g_hi, w.tile, max_delay, N, gl, gh);
            std::vector<int2> pairs;
            for (int b = 0; b < (int)gl.size(); ++b)
                for (int a = 0; a < (int)ql.size(); ++a)
                    if (std::min(qh[a], gh[b]) > std::max(ql[a], gl[b])) pairs.push_back(make_int2(a, b));
            w.n_pairs = (int)pairs.size();
            if (pairs.empty()) pairs.push_back(make_int2(0, 0));
            CK(w.pairs.upload(pairs, c->stream));
        }
        // support-aware work model: 8 T supp_G K  +  8 sum |supp(Q_i) ^ supp((HG)_j)|
        double f = 0;
        for (int j = 0; j < K; ++j) f += 8.0 * c->T * (w.g_hi[j] - w.g_lo[j]);
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
                int a = std::max(w.q_lo[i], w.g_lo[j]), b = std::min(w.q_hi[i], std::min(N, w.g_hi[j] + max_delay));
                if (b > a) f += 8.0 * (b - a);
            }
        w.flops_d = f;
        {   // structural column range of D per 8-row tile: j with supp(Q_i) ^ supp((HG)_j) non-empty for some row i
            const int RT = (K + 7) / 8;
            std::vector<int> jlo(RT, K), jhi(RT, 0);
            double pairs = 0;
            for (int i = 0; i < K; ++i)
                for (int j = 0; j < K; ++j) {
                    int a = std::max(w.q_lo[i], w.g_lo[j]), b = std::min(w.q_hi[i], std::min(N, w.g_hi[j] + max_delay));
                    if (b > a) { jlo[i >> 3] = std::min(jlo[i >> 3], j); jhi[i >> 3] = std::max(jhi[i >> 3], j + 1); if (i != j) pairs += 1; }
                }
            for (int r = 0; r < RT; ++r) if (jhi[r] <= jlo[r]) { jlo[r] = 0; jhi[r] = 0; }
            CK(w.d_jlo.upload(jlo, c->stream)); CK(w.d_jhi.upload(jhi, c->stream));
            w.d_struct_pairs = pairs;
        }
        double fq = 0, fg = 0;
        for (int i = 0; i < K; ++i) { fq += 8.0 * (w.q_hi[i] - w.q_lo[i]); fg += 8.0 * (w.g_hi[i] - w.g_lo[i]); }
        w.flops_demod = fq; w.flops_mod = fg;
    }
    CK(cudaStreamSynchronize(c->stream));
    for (int wfi = 0; wfi < 2; ++wfi) { c->wf[wfi].nsch = 0; }
    for (int si = 0; si < 3; ++si) {
        Scheme& s = c->sch[si];
        if (!s.set) continue;
        if ((!s.mm[0].set || !s.mm[1].set) && !modem_only) return fail(CHEST_ERR_STATE, "scheme without both MMSE variants");
        Waveform& w = c->wf[s.waveform];
        if (w.nsch >= 2) return fail(CHEST_ERR_STATE, "more than two schemes on one waveform");
        if (w.nsch == 1 && c->sch[w.sch[0]].P != s.P) return fail(CHEST_ERR_STATE, "schemes of one waveform must share the pilot count");
        w.sch[w.nsch++] = si;
        c->K_max = std::max(c->K_max, s.K);
        CK(s.xP.alloc((size_t)B * s.P)); CK(s.txword.alloc((size_t)B * s.n_data)); CK(s.txw_t.alloc((size_t)((B + 15) / 16) * 16 * s.n_data)); CK(s.bits.alloc((size_t)B * s.n_bits));
        const size_t S1 = (size_t)std::max(S, 1);
        CK(s.hP.alloc(S1 * B * s.P)); CK(s.hdiag.alloc(S1 * B * s.K));
        CK(s.xD[0].alloc(S1 * B * s.n_data)); CK(s.xD[1].alloc(S1 * B * s.n_data));
    }
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.set && !w.modem_set) continue;
        c->K_max = std::max(c->K_max, w.K);
        int ns = std::max(w.nsch, 1);
        CK(w.x.alloc((size_t)ns * B * w.K)); CK(w.s.alloc((size_t)ns * B * N)); CK(w.r0.alloc((size_t)ns * B * N));
        CK(w.y.alloc((size_t)ns * std::max(S, 1) * B * w.K)); CK(w.htrue.alloc((size_t)B * w.K));
        // D and the H*G planes (dense mode, chest_transmission_matrix) are allocated on first use: ensure_d_buffers
        w.d_alloc_batch = 0;
        if (w.nsch) CK(c->pilot_idx[wfi].alloc((size_t)B * c->sch[w.sch[0]].P));
    }
    CK(c->doppler_u.alloc((size_t)B * c->T * c->paths)); CK(c->phase_u.alloc((size_t)B * c->T * c->paths));
    CK(c->chan_gauss.alloc((size_t)B * (2 * c->n_shift + 1) * c->T));
    CK(c->noise.alloc((size_t)B * std::max(S, 1) * N)); CK(c->h.alloc((size_t)B * c->T * N));
    CK(c->err.alloc((size_t)B * std::max(S, 1) * 17 * 12));
    c->max_batch = B; c->cur_batch = B; c->ctas_for_batch = -1;
    c->finalized = true;
    return CHEST_OK;
}

// ---------------------------------------------------------------- tier 1
int chest_new_realization(uint64_t handle, int batch, const double* du, const double* pu) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(batch >= 1 && batch <= c->max_batch && du && pu);
    if (!(c->fD > 0)) return fail(CHEST_ERR_STATE, "time-invariant channel (f_D = 0): upload the impulse response with chest_set_impulse_response");
    if (c->n_shift > 0) return fail(CHEST_ERR_STATE, "discrete Doppler spectrum: realizations are drawn from normals (chest_new_realization_gauss)");
    CK(cudaSetDevice(c->device));
    size_t n = (size_t)batch * c->T * c->paths;
    CK(cudaMemcpyAsync(c->doppler_u.p, du, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->phase_u.p, pu, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->cur_batch = batch;
    rc = stage_channel(c, batch, c->doppler_u.p, c->phase_u.p); if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

int chest_new_realization_seeded(uint64_t handle, int batch, uint64_t seed, int64_t first_rep) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(batch >= 1 && batch <= c->max_batch);
    if (!(c->fD > 0)) return fail(CHEST_ERR_STATE, "time-invariant channel (f_D = 0): upload the impulse response with chest_set_impulse_response");
    CK(cudaSetDevice(c->device));
    if (c->n_shift > 0) {
        c->cur_batch = batch;
        rc = gen_chan_gauss(c, batch, seed, first_rep); if (rc) return rc;
        rc = stage_channel_discrete(c, batch, c->chan_gauss.p); if (rc) return rc;
        CK(cudaStreamSynchronize(c->stream));
        return CHEST_OK;
    }
    int n = c->T * c->paths;
    dim3 grid((n / 2 + 1 + 127) / 128, batch);
    k_rng_uniform<<<grid, 128, 0, c->stream>>>(c->doppler_u.p, n, batch, RS_DOPPLER, seed, first_rep);
    k_rng_uniform<<<grid, 128, 0, c->stream>>>(c->phase_u.p, n, batch, RS_PHASE, seed, first_rep);
    c->launches += 2;
    c->cur_batch = batch;
    rc = stage_channel(c, batch, c->doppler_u.p, c->phase_u.p); if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

int chest_new_realization_gauss(uint64_t handle, int batch, const double* gauss) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(batch >= 1 && batch <= c->max_batch && gauss);
    if (c->n_shift <= 0) return fail(CHEST_ERR_STATE, "chest_new_realization_gauss needs a 'Discrete-*' Doppler model with f_D > 0");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->chan_gauss.p, gauss, sizeof(cplx) * (size_t)batch * (2 * c->n_shift + 1) * c->T, cudaMemcpyHostToDevice, c->stream));
    c->cur_batch = batch;
    rc = stage_channel_discrete(c, batch, c->chan_gauss.p); if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

int chest_channel_info(uint64_t handle, int* n_doppler_shifts, double* max_doppler_hz, int* n_nonzero_taps) {
    Ctx* c = from(handle);
    ARG(c && c->chan_set);
    if (n_doppler_shifts) *n_doppler_shifts = c->n_shift;
    if (max_doppler_hz) *max_doppler_hz = c->fD;
    if (n_nonzero_taps) *n_nonzero_taps = c->T;
    return CHEST_OK;
}

int chest_set_impulse_response(uint64_t handle, int batch, const double* h) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(batch >= 1 && batch <= c->max_batch && h);
    CK(cudaSetDevice(c->device));
    const size_t N = c->N;
    const cplx* src = reinterpret_cast<const cplx*>(h);
    for (int b = 0; b < batch; ++b)
        for (int t = 0; t < c->T; ++t)
            CK(cudaMemcpyAsync(c->h.p + ((size_t)b * c->T + t) * N, src + ((size_t)b * c->Lt + c->tap_delay[t]) * N,
                               sizeof(cplx) * N, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->cur_batch = batch;
    return CHEST_OK;
}

int chest_get_impulse_response(uint64_t handle, int b, double* out) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(b >= 0 && b < c->cur_batch && out);
    CK(cudaSetDevice(c->device));
    std::vector<cplx> h((size_t)c->T * c->N);
    CK(cudaMemcpyAsync(h.data(), c->h.p + (size_t)b * c->T * c->N, h.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memset(out, 0, sizeof(double) * 2 * (size_t)c->N * c->Lt);          // zero columns for empty taps (FF.m:201)
    for (int t = 0; t < c->T; ++t)
        std::memcpy(out + 2 * (size_t)c->N * c->tap_delay[t], h.data() + (size_t)t * c->N, sizeof(cplx) * c->N);
    return CHEST_OK;
}

int chest_get_convolution_csc(uint64_t handle, int b, int64_t* nnz, int64_t* jc, int32_t* ir, double* val) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(b >= 0 && b < c->cur_batch && nnz);
    const int N = c->N;
    int64_t tot = 0;
    for (int d : c->tap_delay) tot += N - d;
    *nnz = tot;
    if (!jc || !ir || !val) return CHEST_OK;
    CK(cudaSetDevice(c->device));
    std::vector<cplx> h((size_t)c->T * N);
    CK(cudaMemcpyAsync(h.data(), c->h.p + (size_t)b * c->T * N, h.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    // column col holds H[col + d, col] = h[col + d, d] for every non-zero tap d, rows ascending (FF.m:284)
    int64_t e = 0;
    for (int col = 0; col < N; ++col) {
        jc[col] = e;
        for (int t = 0; t < c->T; ++t) {
            int r = col + c->tap_delay[t];
            if (r >= N) continue;
            ir[e] = r; val[2 * e] = h[(size_t)t * N + r].x; val[2 * e + 1] = h[(size_t)t * N + r].y; ++e;
        }
    }
    jc[N] = e;
    return CHEST_OK;
}

int chest_convolve(uint64_t handle, int b, const double* s, int n_cols, double* r) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(b >= 0 && b < c->cur_batch && s && r && n_cols >= 1);
    CK(cudaSetDevice(c->device));
    size_t n = (size_t)n_cols * c->N;
    CK(c->tmp_a.upload(reinterpret_cast<const cplx*>(s), n, c->stream));
    CK(c->tmp_b.alloc(n));
    dim3 grid((c->N + 127) / 128, n_cols);
    k_apply_h<<<grid, 128, 0, c->stream>>>(c->tmp_b.p, c->tmp_a.p, c->h.p, c->d_tap_delay.p, c->N, c->T, 1, b);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(r, c->tmp_b.p, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

int chest_transmission_matrix(uint64_t handle, int b, int wfi, double* D_out, double* h_out) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(b >= 0 && b < c->cur_batch && (wfi == 0 || wfi == 1) && c->wf[wfi].set && D_out);
    CK(cudaSetDevice(c->device));
    Waveform& w = c->wf[wfi];
    rc = stage_transmission_matrix(c, wfi, 1, b); if (rc) return rc;
    const size_t K = w.K, RT8 = ((K + 7) / 8) * 8;
    std::vector<cplx> tiled(RT8 * K);
    CK(cudaMemcpyAsync(tiled.data(), w.D.p + (size_t)b * RT8 * K, sizeof(cplx) * RT8 * K, cudaMemcpyDeviceToHost, c->stream));
    if (h_out) CK(cudaMemcpyAsync(h_out, w.htrue.p + (size_t)b * w.K, sizeof(cplx) * w.K, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cplx* Dh = reinterpret_cast<cplx*>(D_out);                  // device layout [rt][col][8 rows] -> column-major K x K
    for (size_t j = 0; j < K; ++j)
        for (size_t i = 0; i < K; ++i) Dh[i + K * j] = tiled[((i >> 3) * K + j) * 8 + (i & 7)];
    return CHEST_OK;
}

int chest_transmission_matrix_batch(uint64_t handle, int n_rep, int wfi, float* ms, double* flops_per_realization, double* h_out) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(n_rep >= 1 && n_rep <= c->cur_batch && (wfi == 0 || wfi == 1) && c->wf[wfi].set);
    CK(cudaSetDevice(c->device));
    const bool prof = c->profiling;
    c->profiling = true;                                        // the stage records its own events (k_apply_hg / k_gemm_d)
    if (n_rep == 1) CK(cudaEventRecord(c->ev_hg[2 * wfi], c->stream));
    rc = stage_transmission_matrix(c, wfi, n_rep, 0);
    if (n_rep == 1) { CK(cudaEventRecord(c->ev_hg[2 * wfi + 1], c->stream)); CK(cudaEventRecord(c->ev_gd[wfi], c->stream)); }
    c->profiling = prof;
    if (rc) return rc;
    if (h_out) CK(cudaMemcpyAsync(h_out, c->wf[wfi].htrue.p, sizeof(cplx) * (size_t)n_rep * c->wf[wfi].K, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (ms) {
        if (n_rep > 1) { CK(cudaEventElapsedTime(&ms[0], c->ev_hg[2 * wfi], c->ev_hg[2 * wfi + 1])); CK(cudaEventElapsedTime(&ms[1], c->ev_hg[2 * wfi + 1], c->ev_gd[wfi])); }
        else { CK(cudaEventElapsedTime(&ms[1], c->ev_hg[2 * wfi], c->ev_gd[wfi])); ms[0] = 0; }
    }
    if (flops_per_realization) *flops_per_realization = c->wf[wfi].flops_d;
    return CHEST_OK;
}

int chest_transmission_matrix_entries(uint64_t handle, int b, int wfi, int n, const int32_t* rows, const int32_t* cols, double* out) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(b >= 0 && b < c->cur_batch && (wfi == 0 || wfi == 1) && c->wf[wfi].set && n >= 0 && rows && cols && out);
    Waveform& w = c->wf[wfi];
    if (!w.D.p) return fail(CHEST_ERR_STATE, "no transmission matrix on the device: call chest_transmission_matrix_batch first");
    CK(cudaSetDevice(c->device));
    const size_t K = w.K, RT8 = ((K + 7) / 8) * 8;
    for (int e = 0; e < n; ++e) {
        ARG(rows[e] >= 0 && rows[e] < (int)K && cols[e] >= 0 && cols[e] < (int)K);
        const size_t i = rows[e], j = cols[e];
        CK(cudaMemcpyAsync(out + 2 * e, w.D.p + (size_t)b * RT8 * K + ((i >> 3) * K + j) * 8 + (i & 7), sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

static int plain_gemm(Ctx* c, int wfi, bool demod, const double* in, int n_cols, double* out) {
    Waveform& w = c->wf[wfi];
    const int N = c->N, K = w.K;
    const int len_in = demod ? N : K, len_out = demod ? K : N;
    CK(c->tmp_a.upload(reinterpret_cast<const cplx*>(in), (size_t)n_cols * len_in, c->stream));
    CK(c->tmp_b.alloc((size_t)n_cols * len_out));
    GemmParams p{};
    p.M = len_out; p.Kc = len_in; p.n_cols = n_cols; p.ldc = len_out; p.out = c->tmp_b.p;
    p.bsrc = c->tmp_a.p; p.ldb = len_in;
    if (demod) { p.At = w.Q.p; p.lda = N; p.conj_a = 1; p.mt_klo = w.q_klo.p; p.mt_khi = w.q_khi.p; }
    else { p.At = w.Gt.p; p.lda = K; p.conj_a = 0; p.mt_klo = w.gt_klo.p; p.mt_khi = w.gt_khi.p; }
    CK(launch_gemm<GEMM_PLAIN>(c, p, 1, w.tile));
    CK(cudaMemcpyAsync(out, c->tmp_b.p, sizeof(cplx) * (size_t)n_cols * len_out, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}
int chest_modulate(uint64_t handle, int wfi, const double* x, int n_cols, double* s) {
    Ctx* c = from(handle);
    ARG(c && (wfi == 0 || wfi == 1) && c->wf[wfi].set && x && s && n_cols >= 1);
    CK(cudaSetDevice(c->device));
    return plain_gemm(c, wfi, false, x, n_cols, s);
}
int chest_demodulate(uint64_t handle, int wfi, const double* r, int n_cols, double* y) {
    Ctx* c = from(handle);
    ARG(c && (wfi == 0 || wfi == 1) && c->wf[wfi].set && r && y && n_cols >= 1);
    CK(cudaSetDevice(c->device));
    return plain_gemm(c, wfi, true, r, n_cols, y);
}

// ---------------------------------------------------------------- FFT modem
int chest_set_modem(uint64_t handle, int wfi, int kind, int L, int Ksym, int nfft, const int32_t* bin, int time_spacing,
                    int O, int cp, int zero_guard, const double* filt, const double* phase, double norm, double F) {
    Ctx* c = from(handle);
    ARG(c && (wfi == 0 || wfi == 1) && (kind == 0 || kind == 1) && L >= 1 && Ksym >= 1 && nfft >= 2 && nfft <= 4096 && bin);
    ARG(L <= nfft && time_spacing >= 1 && norm > 0 && F > 0);
    if (kind == 0) ARG(O >= 1 && filt && phase);
    else ARG(cp >= 0 && zero_guard >= 0 && time_spacing == nfft + cp);
    CK(cudaSetDevice(c->device));
    Waveform& w = c->wf[wfi];
    ModemDev& md = w.modem;
    md = ModemDev{};
    md.kind = kind; md.L = L; md.Ksym = Ksym; md.nfft = nfft; md.time_spacing = time_spacing; md.O = O; md.cp = cp;
    md.zero_guard = zero_guard; md.Np = kind == 0 ? O * nfft : 0;
    md.N = kind == 0 ? md.Np + (Ksym - 1) * time_spacing : 2 * zero_guard + Ksym * time_spacing;
    if (c->N && c->N != md.N) return fail(CHEST_ERR_ARG, "Total number of samples must be the same for the channel and every waveform");
    md.norm = norm; md.inv_demod = kind == 0 ? 1.0 / (norm * F) : 1.0 / norm;
    md.plan.n = nfft; md.plan.n_stage = 0;
    int rem = nfft;
    for (int r : {4, 2, 3, 5, 7, 11, 13})
        while (rem % r == 0) { ARG(md.plan.n_stage < FFT_MAX_STAGES); md.plan.radix[md.plan.n_stage++] = r; rem /= r; }
    if (rem != 1) return fail(CHEST_ERR_ARG, "FFT size has a prime factor above 13");
    std::vector<int> b(bin, bin + L);
    std::vector<char> used(nfft, 0);
    for (int x : b) { ARG(x >= 0 && x < nfft && !used[x]); used[x] = 1; }
    std::vector<cplx> tw(nfft);
    for (int m = 0; m < nfft; ++m) {                            // exact octant symmetry is not needed: cos/sin of 2 pi m / n in double
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * m / nfft;
        tw[m] = cmake((double)cosl(a), (double)sinl(a));
    }
    cudaStream_t st = c->stream;
    CK(w.m_bin.upload(b, st)); CK(w.m_tw.upload(tw, st));
    if (kind == 0) {
        CK(w.m_filt.upload(filt, (size_t)md.Np, st));
        CK(w.m_phase.upload(reinterpret_cast<const cplx*>(phase), (size_t)L * Ksym, st));
    }
    CK(cudaStreamSynchronize(st));
    md.bin = w.m_bin.p; md.filt = w.m_filt.p; md.phase = w.m_phase.p; md.tw = w.m_tw.p;
    const size_t smem = (size_t)3 * nfft * sizeof(cplx);
    CK(cudaFuncSetAttribute(k_modem_ifft, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_modem_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ARG(smem <= 200 * 1024);
    if (!w.set) { w.K = L * Ksym; c->N = md.N; c->finalized = false; }       // modem-only waveform (no dense G / Q)
    w.modem_set = true; w.pf_state = 0; c->ctas_for_batch = -1;
    return CHEST_OK;
}

int chest_modulate_fft(uint64_t handle, int wfi, const double* x, int n_cols, double* s_out) {
    Ctx* c = from(handle);
    ARG(c && (wfi == 0 || wfi == 1) && c->wf[wfi].modem_set && x && s_out && n_cols >= 1);
    CK(cudaSetDevice(c->device));
    Waveform& w = c->wf[wfi];
    const ModemDev& md = w.modem;
    const size_t nx = (size_t)n_cols * md.L * md.Ksym, ns = (size_t)n_cols * md.N;
    CK(w.m_in.upload(reinterpret_cast<const cplx*>(x), nx, c->stream));
    CK(w.m_out.alloc(ns));
    int rc = modem_modulate_dev(c, w, w.m_in.p, n_cols, w.m_out.p); if (rc) return rc;
    CK(cudaMemcpyAsync(s_out, w.m_out.p, ns * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}
int chest_demodulate_fft(uint64_t handle, int wfi, const double* r, int n_cols, double* y_out) {
    Ctx* c = from(handle);
    ARG(c && (wfi == 0 || wfi == 1) && c->wf[wfi].modem_set && r && y_out && n_cols >= 1);
    CK(cudaSetDevice(c->device));
    Waveform& w = c->wf[wfi];
    const ModemDev& md = w.modem;
    const size_t ny = (size_t)n_cols * md.L * md.Ksym, nr = (size_t)n_cols * md.N;
    CK(w.m_in.upload(reinterpret_cast<const cplx*>(r), nr, c->stream));
    CK(w.m_out.alloc(ny));
    int rc = modem_demodulate_dev(c, w, w.m_in.p, n_cols, w.m_out.p); if (rc) return rc;
    CK(cudaMemcpyAsync(y_out, w.m_out.p, ny * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

// ---------------------------------------------------------------- SimpleVersion_DoublyFlat.m chain
int chest_set_interpolation(uint64_t handle, int si, const double* M) {
    Ctx* c = from(handle);
    ARG(c && si >= 0 && si < 3 && c->sch[si].set && M);
    CK(cudaSetDevice(c->device));
    Scheme& s = c->sch[si];
    const cplx* m = reinterpret_cast<const cplx*>(M);          // K x P column-major -> row-major
    std::vector<cplx> rm((size_t)s.K * s.P);
    for (int i = 0; i < s.K; ++i) for (int p = 0; p < s.P; ++p) rm[(size_t)i * s.P + p] = m[(size_t)i + (size_t)s.K * p];
    CK(s.interp.upload(rm, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

int chest_sv_run_batch(uint64_t handle, int n_body, const double* pn_time, const chest_sv_draws* draws, uint64_t seed,
                       int64_t first_body, uint32_t* err_out) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(n_body >= 1 && n_body <= c->max_batch && pn_time && err_out);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int N = c->N;
    for (int wfi = 0; wfi < 2; ++wfi) if (c->wf[wfi].nsch && !c->wf[wfi].modem_set) return fail(CHEST_ERR_STATE, "chest_sv_run_batch needs chest_set_modem for every waveform in use");
    for (int si = 0; si < 3; ++si) if (c->sch[si].set && !c->sch[si].interp.p) return fail(CHEST_ERR_STATE, "chest_sv_run_batch needs chest_set_interpolation for every scheme");
    c->cur_batch = n_body;
    CK(c->sv_h.alloc(c->max_batch)); CK(c->sv_noise.alloc((size_t)c->max_batch * 2 * N)); CK(c->sv_pn.alloc(c->max_batch));
    CK(c->sv_err.alloc((size_t)c->max_batch * 5));
    CK(cudaMemcpyAsync(c->sv_pn.p, pn_time, sizeof(double) * n_body, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(c->sv_err.p, 0, sizeof(uint32_t) * n_body * 5, st));
    const uint8_t* bits[3] = {c->sch[0].bits.p, c->sch[1].bits.p, c->sch[2].bits.p};
    const int32_t* pidx[2] = {c->pilot_idx[0].p, c->pilot_idx[1].p};
    if (draws) {                                                   // explicit draws (host pointers), SV.m:95-128 order
        ARG(!draws->on_device && draws->h && draws->noise[0] && draws->noise[1]);
        CK(cudaMemcpyAsync(c->sv_h.p, draws->h, sizeof(cplx) * n_body, cudaMemcpyHostToDevice, st));
        for (int wfi = 0; wfi < 2; ++wfi)                          // host layout per waveform [body][N] -> device [body][wf][N]
            CK(cudaMemcpy2DAsync(c->sv_noise.p + (size_t)wfi * N, sizeof(cplx) * 2 * N, draws->noise[wfi], sizeof(cplx) * N,
                                 sizeof(cplx) * N, n_body, cudaMemcpyHostToDevice, st));
        for (int i = 0; i < 3; ++i)
            if (c->sch[i].set) { ARG(draws->bits[i]); CK(cudaMemcpyAsync(c->sch[i].bits.p, draws->bits[i], (size_t)n_body * c->sch[i].n_bits, cudaMemcpyHostToDevice, st)); }
        for (int i = 0; i < 2; ++i)
            if (c->wf[i].nsch) { ARG(draws->pilot_idx[i]); CK(cudaMemcpyAsync(c->pilot_idx[i].p, draws->pilot_idx[i], sizeof(int32_t) * n_body * c->sch[c->wf[i].sch[0]].P, cudaMemcpyHostToDevice, st)); }
    } else {
        for (int si = 0; si < 3; ++si) {
            Scheme& s = c->sch[si];
            if (!s.set) continue;
            dim3 gb(((s.n_bits + 127) / 128 + 127) / 128, n_body);
            k_rng_bits<<<gb, 128, 0, st>>>(s.bits.p, s.n_bits, n_body, RS_BITS0 + si, seed, first_body);
            c->launches++;
        }
        for (int wfi = 0; wfi < 2; ++wfi) {
            Waveform& w = c->wf[wfi];
            if (!w.nsch) continue;
            Scheme& s = c->sch[w.sch[0]];
            dim3 gp((s.P + 63) / 64, n_body);
            k_rng_index<<<gp, 64, 0, st>>>(c->pilot_idx[wfi].p, s.P, n_body, c->cst[s.constellation].order, RS_PILOT0 + wfi, seed, first_body);
            c->launches++;
        }
        dim3 gh(1, n_body);
        k_rng_cnormal<<<gh, 32, 0, st>>>(c->sv_h.p, 1, n_body, RS_SV_H, seed, first_body, 0.70710678118654752440);   // h ~ CN(0,1), SV.m:123
        dim3 gn((N + 127) / 128, 2, n_body);
        k_rng_normal<<<gn, 128, 0, st>>>(c->sv_noise.p, N, 2, n_body, seed, first_body);
        c->launches += 2;
        CK(cudaGetLastError());
    }
    for (int si = 0; si < 3; ++si) {
        if (!c->sch[si].set) continue;
        SchemeDev sd = scheme_dev(c, si);
        k_tx_symbols<<<n_body, 256, sd.K_in * sizeof(cplx), st>>>(sd, c->cst[sd.constellation].dev, bits[si], pidx[sd.waveform], n_body);
        c->launches++;
        CK(cudaGetLastError());
    }
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.nsch) continue;
        const int n_cols = w.nsch * n_body;
        rc = modem_modulate_dev(c, w, w.x.p, n_cols, w.s.p); if (rc) return rc;                       // SV.m:118-120
        dim3 g(n_cols, (N + 127) / 128);
        k_sv_channel<<<g, 128, 0, st>>>(w.r0.p, w.s.p, c->sv_h.p, c->sv_noise.p, c->sv_pn.p, N, n_body, wfi);   // SV.m:123-131
        c->launches++;
        CK(cudaGetLastError());
        rc = modem_demodulate_dev(c, w, w.r0.p, n_cols, w.y.p); if (rc) return rc;                    // SV.m:133-135
    }
    // error slots: 0 aux, 1 cod, 2 FBMC perfect (data spreading with the true channel), 3 OFDM, 4 OFDM perfect (SV.m:165-169)
    const int slot_est[3] = {0, 1, 3}, slot_perf[3] = {-1, 2, 4};
    for (int si = 0; si < 3; ++si) {
        if (!c->sch[si].set) continue;
        SchemeDev sd = scheme_dev(c, si);
        sd.y = c->wf[sd.waveform].y.p + (size_t)((c->wf[sd.waveform].sch[0] == si) ? 0 : 1) * n_body * sd.K;   // one "SNR point": [g][body][K]
        const size_t smem = ((size_t)sd.P + 2 * sd.K) * sizeof(cplx);
        k_sv_detect<<<n_body, 256, smem, st>>>(sd, c->cst[sd.constellation].dev, sd.y, c->sch[si].interp.p, c->sv_h.p, c->sv_err.p,
                                                 slot_est[si], slot_perf[si], n_body);
        c->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(err_out, c->sv_err.p, sizeof(uint32_t) * n_body * 5, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CHEST_OK;
}

int chest_estimate(uint64_t handle, int si, int variant, int i_snr, const double* hP, double* Dhat_out, double* hdiag_out) {
    Ctx* c = from(handle);
    ARG(c && si >= 0 && si < 3 && (variant == 0 || variant == 1) && hP);
    Scheme& s = c->sch[si];
    ARG(s.set && s.mm[variant].set && i_snr >= 0 && i_snr < c->S);
    CK(cudaSetDevice(c->device));
    MmseVariant& m = s.mm[variant];
    const int K = s.K, P = s.P;
    CK(c->tmp_a.upload(reinterpret_cast<const cplx*>(hP), P, c->stream));
    CK(c->tmp_b.alloc((size_t)K * K + K));
    CK(cudaMemsetAsync(c->tmp_b.p, 0, sizeof(cplx) * ((size_t)K * K + K), c->stream));
    WTiles wt; wt.frag = m.frag[i_snr].p; wt.diag = m.diag[i_snr].p;
    int n = std::max(m.n_tiles * 8, K);
    k_estimate<<<(n + 255) / 256, 256, 0, c->stream>>>(c->tmp_b.p, c->tmp_b.p + (size_t)K * K, wt, m.tile_ptr.p,
                                                       m.tile_delta.p, c->tmp_a.p, K, P, (P + 3) / 4, (K + 7) / 8);
    c->launches++;
    CK(cudaGetLastError());
    if (Dhat_out) CK(cudaMemcpyAsync(Dhat_out, c->tmp_b.p, sizeof(cplx) * (size_t)K * K, cudaMemcpyDeviceToHost, c->stream));
    if (hdiag_out) CK(cudaMemcpyAsync(hdiag_out, c->tmp_b.p + (size_t)K * K, sizeof(cplx) * K, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

// ---------------------------------------------------------------- tier 2
int chest_prefetch_draws(uint64_t handle, int n_rep, const chest_draws* host, chest_draws* dev) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(n_rep >= 1 && n_rep <= c->max_batch && host && dev && !host->on_device);
    const bool discrete = c->n_shift > 0;
    const size_t n_cg = (size_t)(2 * c->n_shift + 1) * c->T;
    ARG(host->noise && (discrete ? host->channel_gauss != nullptr : (host->doppler_u && host->phase_u)));
    CK(cudaSetDevice(c->device));
    const int S = c->S, N = c->N, TP = c->T * c->paths, B = c->max_batch;
    for (auto& a : c->pf) {                                        // both sets are allocated by the first call
        CK(a.du.alloc((size_t)B * TP)); CK(a.pu.alloc((size_t)B * TP)); CK(a.noise.alloc((size_t)B * S * N));
        CK(a.cg.alloc((size_t)B * n_cg));
        for (int i = 0; i < 3; ++i) if (c->sch[i].set) CK(a.bits[i].alloc((size_t)B * c->sch[i].n_bits));
        for (int i = 0; i < 2; ++i) if (c->wf[i].set && c->wf[i].nsch) CK(a.pidx[i].alloc((size_t)B * c->sch[c->wf[i].sch[0]].P));
    }
    Ctx::Prefetch& q = c->pf[c->pf_next];
    c->pf_next ^= 1;
    cudaStream_t cs = c->copy_stream;
    if (q.used) CK(cudaStreamWaitEvent(cs, q.released, 0));       // the batch that read this set has finished with it
    if (discrete) CK(cudaMemcpyAsync(q.cg.p, host->channel_gauss, sizeof(cplx) * n_rep * n_cg, cudaMemcpyHostToDevice, cs));
    else {
        CK(cudaMemcpyAsync(q.du.p, host->doppler_u, sizeof(double) * n_rep * TP, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(q.pu.p, host->phase_u, sizeof(double) * n_rep * TP, cudaMemcpyHostToDevice, cs));
    }
    CK(cudaMemcpyAsync(q.noise.p, host->noise, sizeof(cplx) * (size_t)n_rep * S * N, cudaMemcpyHostToDevice, cs));
    *dev = chest_draws{};
    dev->channel_gauss = discrete ? reinterpret_cast<const double*>(q.cg.p) : nullptr;
    for (int i = 0; i < 3; ++i)
        if (c->sch[i].set) {
            ARG(host->bits[i]);
            CK(cudaMemcpyAsync(q.bits[i].p, host->bits[i], (size_t)n_rep * c->sch[i].n_bits, cudaMemcpyHostToDevice, cs));
            dev->bits[i] = q.bits[i].p;
        }
    for (int i = 0; i < 2; ++i)
        if (c->wf[i].set && c->wf[i].nsch) {
            ARG(host->pilot_idx[i]);
            const int P = c->sch[c->wf[i].sch[0]].P;
            CK(cudaMemcpyAsync(q.pidx[i].p, host->pilot_idx[i], sizeof(int32_t) * n_rep * P, cudaMemcpyHostToDevice, cs));
            dev->pilot_idx[i] = q.pidx[i].p;
        }
    CK(cudaEventRecord(q.landed, cs));
    q.used = true;
    dev->doppler_u = q.du.p; dev->phase_u = q.pu.p; dev->noise = reinterpret_cast<const double*>(q.noise.p);
    dev->on_device = 1;
    return CHEST_OK;
}

int64_t chest_draws_bytes(uint64_t handle, int n_rep) {
    Ctx* c = from(handle);
    if (!c) return 0;
    int64_t per = (c->n_shift > 0 ? (int64_t)(2 * c->n_shift + 1) * c->T * 16 : 2 * (int64_t)c->T * c->paths * 8) + (int64_t)c->S * c->N * 16;
    for (int si = 0; si < 3; ++si) if (c->sch[si].set) per += c->sch[si].n_bits;
    for (int wfi = 0; wfi < 2; ++wfi) if (c->wf[wfi].set && c->wf[wfi].nsch) per += 4 * c->sch[c->wf[wfi].sch[0]].P;
    return per * n_rep;
}

int chest_generate_draws(uint64_t handle, int n_rep, uint64_t seed, int64_t first_rep, chest_draws* out) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(n_rep >= 1 && n_rep <= c->max_batch && out);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    int n = c->T * c->paths;
    if (c->n_shift > 0) { int rc2 = gen_chan_gauss(c, n_rep, seed, first_rep); if (rc2) return rc2; }
    else {
        dim3 gu((n / 2 + 1 + 127) / 128, n_rep);
        k_rng_uniform<<<gu, 128, 0, st>>>(c->doppler_u.p, n, n_rep, RS_DOPPLER, seed, first_rep);
        k_rng_uniform<<<gu, 128, 0, st>>>(c->phase_u.p, n, n_rep, RS_PHASE, seed, first_rep);
        c->launches += 2;
    }
    std::memset(out, 0, sizeof(*out));
    out->channel_gauss = c->n_shift > 0 ? reinterpret_cast<const double*>(c->chan_gauss.p) : nullptr;
    for (int si = 0; si < 3; ++si) {
        Scheme& s = c->sch[si];
        if (!s.set) continue;
        dim3 gb(((s.n_bits + 127) / 128 + 127) / 128, n_rep);
        k_rng_bits<<<gb, 128, 0, st>>>(s.bits.p, s.n_bits, n_rep, RS_BITS0 + si, seed, first_rep);
        c->launches++;
        out->bits[si] = s.bits.p;
    }
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.set || !w.nsch) continue;
        Scheme& s = c->sch[w.sch[0]];
        dim3 gp((s.P + 63) / 64, n_rep);
        k_rng_index<<<gp, 64, 0, st>>>(c->pilot_idx[wfi].p, s.P, n_rep, c->cst[s.constellation].order, RS_PILOT0 + wfi, seed, first_rep);
        c->launches++;
        out->pilot_idx[wfi] = c->pilot_idx[wfi].p;
    }
    dim3 gn((c->N + 127) / 128, c->S, n_rep);
    k_rng_normal<<<gn, 128, 0, st>>>(c->noise.p, c->N, c->S, n_rep, seed, first_rep);
    c->launches++;
    CK(cudaGetLastError());
    out->doppler_u = c->doppler_u.p; out->phase_u = c->phase_u.p;
    out->noise = reinterpret_cast<const double*>(c->noise.p); out->on_device = 1;
    return CHEST_OK;
}

int chest_download_draws(uint64_t handle, int n_rep, double* du, double* pu, uint8_t* b0, uint8_t* b1, uint8_t* b2,
                         int32_t* p0, int32_t* p1, double* noise) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(n_rep >= 1 && n_rep <= c->max_batch);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    size_t n = (size_t)n_rep * c->T * c->paths;
    if (du) CK(cudaMemcpyAsync(du, c->doppler_u.p, n * 8, cudaMemcpyDeviceToHost, st));
    if (pu) CK(cudaMemcpyAsync(pu, c->phase_u.p, n * 8, cudaMemcpyDeviceToHost, st));
    uint8_t* bb[3] = {b0, b1, b2};
    for (int si = 0; si < 3; ++si)
        if (bb[si] && c->sch[si].set) CK(cudaMemcpyAsync(bb[si], c->sch[si].bits.p, (size_t)n_rep * c->sch[si].n_bits, cudaMemcpyDeviceToHost, st));
    int32_t* pp[2] = {p0, p1};
    for (int wfi = 0; wfi < 2; ++wfi)
        if (pp[wfi] && c->wf[wfi].nsch) CK(cudaMemcpyAsync(pp[wfi], c->pilot_idx[wfi].p, sizeof(int32_t) * n_rep * c->sch[c->wf[wfi].sch[0]].P, cudaMemcpyDeviceToHost, st));
    if (noise) CK(cudaMemcpyAsync(noise, c->noise.p, sizeof(cplx) * (size_t)n_rep * c->S * c->N, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CHEST_OK;
}

int chest_run_batch(uint64_t handle, int n_rep, int n_iter, const chest_draws* draws, uint64_t seed,
                    int64_t first_rep, uint32_t* err_out) {
    Ctx* c = from(handle);
    ARG(c && err_out);
    CK(cudaSetDevice(c->device));
    return run_pipeline(c, n_rep, n_iter, draws, seed, first_rep, err_out, nullptr);
}
int chest_run_batch_device(uint64_t handle, int n_rep, int n_iter, const chest_draws* draws, uint64_t seed,
                           int64_t first_rep, uint32_t* err_dev) {
    Ctx* c = from(handle);
    ARG(c);
    CK(cudaSetDevice(c->device));
    return run_pipeline(c, n_rep, n_iter, draws, seed, first_rep, nullptr, err_dev);
}

// ---------------------------------------------------------------- asynchronous run / wait, several devices
int chest_run_batch_async(uint64_t handle, int n_rep, int n_iter, const chest_draws* draws, uint64_t seed,
                          int64_t first_rep) {
    Ctx* c = from(handle);
    ARG(c);
    CK(cudaSetDevice(c->device));
    return run_pipeline(c, n_rep, n_iter, draws, seed, first_rep, nullptr, nullptr, false);
}

int chest_wait(uint64_t handle, uint32_t* err_out) {
    Ctx* c = from(handle);
    ARG(c);
    if (!c->pending) return fail(CHEST_ERR_STATE, "no asynchronous run is pending on this context");
    CK(cudaSetDevice(c->device));
    const size_t n = c->pending_n_err;
    int rc = finish_pipeline(c);
    if (rc) return rc;
    if (err_out) std::memcpy(err_out, c->err_pinned, n * sizeof(uint32_t));
    return CHEST_OK;
}

namespace {
// NCCL is bound at run time (dlopen): the library has no link-time dependency on it, and inside a process that has
// already loaded an NCCL (PyTorch's) the same soname resolves to that copy.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& why) {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { why = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        bool ok = true;
        auto sym = [&](const char* n) { void* f = dlsym(lib, n); if (!f) { ok = false; why = std::string("NCCL symbol missing: ") + n; } return f; };
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(lib); lib = nullptr; }
        return ok;
    }
};
NcclApi g_nccl;
#define NCK(call)                                                                                  \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            return fail(CHEST_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r_));    \
    } while (0)

struct Multi {
    std::vector<Ctx*> ctx;
    std::vector<ncclComm_t> comm;                  // empty for a single device
    std::vector<cudaEvent_t> ev0, ev1;             // around the all-reduce, per device
};
Multi* multi_from(uint64_t h) { return reinterpret_cast<Multi*>(static_cast<uintptr_t>(h)); }
}  // namespace

int chest_multi_create(const uint64_t* handles, int n_devices, uint64_t* multi) {
    ARG(handles && multi && n_devices >= 1 && n_devices <= 64);
    Multi* m = new Multi();
    std::vector<int> devs;
    for (int i = 0; i < n_devices; ++i) {
        Ctx* c = from(handles[i]);
        if (!c || !c->finalized) { delete m; return fail(CHEST_ERR_STATE, "chest_multi_create: every context must be finalized"); }
        for (int d : devs) if (d == c->device) { delete m; return fail(CHEST_ERR_ARG, "chest_multi_create: two contexts on one device"); }
        if (i > 0 && (c->S != m->ctx[0]->S || c->N != m->ctx[0]->N)) { delete m; return fail(CHEST_ERR_ARG, "chest_multi_create: contexts differ in configuration"); }
        m->ctx.push_back(c); devs.push_back(c->device);
    }
    if (n_devices > 1) {
        std::string why;
        if (!g_nccl.load(why)) { delete m; return fail(CHEST_ERR_STATE, why); }
        m->comm.resize(n_devices);
        ncclResult_t r = g_nccl.CommInitAll(m->comm.data(), n_devices, devs.data());
        if (r != ncclSuccess) { delete m; return fail(CHEST_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r)); }
    }
    m->ev0.resize(n_devices); m->ev1.resize(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        CK(cudaSetDevice(devs[i]));
        CK(cudaEventCreate(&m->ev0[i])); CK(cudaEventCreate(&m->ev1[i]));
    }
    *multi = (uint64_t)(uintptr_t)m;
    return CHEST_OK;
}

int chest_multi_destroy(uint64_t multi) {
    Multi* m = multi_from(multi);
    if (!m) return CHEST_OK;
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        cudaSetDevice(m->ctx[i]->device);
        cudaEventDestroy(m->ev0[i]); cudaEventDestroy(m->ev1[i]);
    }
    for (auto& q : m->comm) g_nccl.CommDestroy(q);
    delete m;
    return CHEST_OK;
}

int chest_multi_run(uint64_t multi, int64_t n_rep_total, int n_iter, uint64_t seed, int64_t first_rep,
                    uint32_t* err_out, uint64_t* totals_out, float* reduce_ms) {
    Multi* m = multi_from(multi);
    ARG(m && n_rep_total >= 1 && n_iter >= 0 && n_iter <= 16);
    const int nd = (int)m->ctx.size(), S = m->ctx[0]->S;
    const size_t per_rep = (size_t)S * (n_iter + 1) * 12;
    for (Ctx* c : m->ctx) {
        CK(cudaSetDevice(c->device));
        CK(c->totals.alloc(per_rep));
        CK(cudaMemsetAsync(c->totals.p, 0, per_rep * sizeof(unsigned long long), c->stream));
    }
    // contiguous blocks of realizations per device, handed out in rounds of at most max_batch each; all devices of a
    // round are enqueued before any is waited for (the host thread never blocks between them)
    int64_t done = 0;
    std::vector<int64_t> off(nd); std::vector<int> cnt(nd);
    while (done < n_rep_total) {
        for (int i = 0; i < nd; ++i) {
            Ctx* c = m->ctx[i];
            cnt[i] = (int)std::min<int64_t>(c->max_batch, n_rep_total - done);
            off[i] = done; done += cnt[i];
            if (cnt[i] == 0) continue;
            CK(cudaSetDevice(c->device));
            int rc = run_pipeline(c, cnt[i], n_iter, nullptr, seed, first_rep + off[i], nullptr, nullptr, false);
            if (rc) return rc;
            k_sum_counters<<<(unsigned)per_rep, 128, 0, c->stream>>>(c->totals.p, c->err.p, cnt[i], (int)per_rep);
            c->launches++;
            CK(cudaGetLastError());
        }
        for (int i = 0; i < nd; ++i) {
            if (cnt[i] == 0) continue;
            Ctx* c = m->ctx[i];
            CK(cudaSetDevice(c->device));
            const size_t n = c->pending_n_err;
            int rc = finish_pipeline(c);
            if (rc) return rc;
            if (err_out) std::memcpy(err_out + (size_t)off[i] * per_rep, c->err_pinned, n * sizeof(uint32_t));
        }
    }
    // the one collective of the path: sum of the per-GPU counter totals
    for (int i = 0; i < nd; ++i) { CK(cudaSetDevice(m->ctx[i]->device)); CK(cudaEventRecord(m->ev0[i], m->ctx[i]->stream)); }
    if (nd > 1) {
        NCK(g_nccl.GroupStart());
        for (int i = 0; i < nd; ++i)
            NCK(g_nccl.AllReduce(m->ctx[i]->totals.p, m->ctx[i]->totals.p, per_rep, ncclUint64, ncclSum, m->comm[i], m->ctx[i]->stream));
        NCK(g_nccl.GroupEnd());
    }
    float worst = 0;
    for (int i = 0; i < nd; ++i) {
        CK(cudaSetDevice(m->ctx[i]->device));
        CK(cudaEventRecord(m->ev1[i], m->ctx[i]->stream));
        CK(cudaStreamSynchronize(m->ctx[i]->stream));
        float t = 0; CK(cudaEventElapsedTime(&t, m->ev0[i], m->ev1[i]));
        worst = std::max(worst, t);
    }
    if (reduce_ms) *reduce_ms = worst;
    if (totals_out) {
        CK(cudaSetDevice(m->ctx[0]->device));
        CK(cudaMemcpy(totals_out, m->ctx[0]->totals.p, per_rep * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    return CHEST_OK;
}

int chest_bit_counts(uint64_t handle, int64_t* n_bits) {
    Ctx* c = from(handle);
    ARG(c && n_bits);
    for (int si = 0; si < 3; ++si) {
        n_bits[2 * si] = c->sch[si].set ? c->sch[si].n_bits : 0;
        n_bits[2 * si + 1] = c->sch[si].set ? c->sch[si].n_bits_edge : 0;
    }
    return CHEST_OK;
}

int chest_get_state(uint64_t handle, int what, int si, int rep, int i_snr, double* out) {
    Ctx* c = from(handle);
    int rc = check_ready(c); if (rc) return rc;
    ARG(si >= 0 && si < 3 && c->sch[si].set && rep >= 0 && rep < c->cur_batch && i_snr >= 0 && i_snr < c->S && out);
    CK(cudaSetDevice(c->device));
    Scheme& s = c->sch[si];
    SchemeDev d = scheme_dev(c, si);
    size_t col = (size_t)i_snr * c->cur_batch + rep;
    const cplx* src = nullptr; size_t n = 0;
    switch (what) {
        case 0: src = d.y + col * s.K; n = s.K; break;
        case 1: src = d.hP + col * s.P; n = s.P; break;
        case 2: src = d.xD[0] + col * s.n_data; n = s.n_data; break;
        case 3: src = d.xD[1] + col * s.n_data; n = s.n_data; break;
        case 4: src = d.hdiag + col * s.K; n = s.K; break;
        default: return fail(CHEST_ERR_ARG, "unknown state selector");
    }
    CK(cudaMemcpyAsync(out, src, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHEST_OK;
}

int64_t chest_launch_count(uint64_t handle) { Ctx* c = from(handle); return c ? c->launches : 0; }

int chest_set_profiling(uint64_t handle, int enable) {
    Ctx* c = from(handle);
    ARG(c);
    c->profiling = enable != 0;
    return CHEST_OK;
}
int chest_stage_times(uint64_t handle, float* ms) {
    Ctx* c = from(handle);
    ARG(c && ms);
    for (int i = 0; i < 7; ++i) ms[i] = c->stage_ms[i];
    return CHEST_OK;
}
int chest_banded_apply_stats(uint64_t handle, float* ms, double* bytes) {
    Ctx* c = from(handle);
    ARG(c && ms && bytes);
    *ms = c->hg_ms; *bytes = c->hg_bytes;
    return CHEST_OK;
}

int chest_set_perfect_csi_mode(uint64_t handle, int mode) {
    Ctx* c = from(handle);
    ARG(c && (mode == 0 || mode == 1));
    c->perf_mode = mode;
    c->ctas_for_batch = -1;                                    // unit order and column tables depend on the mode
    return CHEST_OK;
}

int chest_set_mse_accumulation(uint64_t handle, int enable) {
    Ctx* c = from(handle);
    ARG(c);
    c->mse_on = enable != 0;
    return CHEST_OK;
}

int chest_get_mse(uint64_t handle, double* out) {
    Ctx* c = from(handle);
    ARG(c && out);
    if (!c->mse_on || !c->mse.p) return fail(CHEST_ERR_STATE, "no MSE sums: enable chest_set_mse_accumulation before the run");
    if (c->pending) return fail(CHEST_ERR_STATE, "an asynchronous run is pending on this context: call chest_wait first");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(out, c->mse.p, sizeof(double) * (size_t)c->cur_batch * c->S * (c->last_iter + 1) * 3, cudaMemcpyDeviceToHost));
    return CHEST_OK;
}

int chest_set_precision(uint64_t handle, int mode) {
    Ctx* c = from(handle);
    ARG(c && (mode == CHEST_PRECISION_FP64 || mode == CHEST_PRECISION_SPLIT_BF16));
    if (c->pending) return fail(CHEST_ERR_STATE, "an asynchronous run is pending on this context");
    c->precision = mode;
    c->ctas_for_batch = -1;                                    // the work items of the tensor-core kernel are built with the unit list
    return CHEST_OK;
}

int chest_precision_info(uint64_t handle, int* mode, double* mma_flops_per_launch, int64_t* operand_bytes) {
    Ctx* c = from(handle);
    ARG(c);
    if (mode) *mode = c->precision;
    if (mma_flops_per_launch) *mma_flops_per_launch = c->tc_mma_flops;
    if (operand_bytes) {
        int64_t b = 0;
        for (auto& s : c->sch) for (auto& m : s.mm) b += (int64_t)m.tc_img.n;
        *operand_bytes = b;
    }
    return CHEST_OK;
}

int chest_set_estimator_mode(uint64_t handle, int mode) {
    Ctx* c = from(handle);
    ARG(c && mode >= CHEST_ESTIMATOR_AUTO && mode <= CHEST_ESTIMATOR_FACTORED_EXACT);
    if (c->pending) return fail(CHEST_ERR_STATE, "an asynchronous run is pending on this context");
    c->est_mode = mode;
    c->ctas_for_batch = -1;
    return CHEST_OK;
}

int chest_set_pseudo_channels(uint64_t handle, int wfi, int n_pilots, const double* M, double removed_max) {
    Ctx* c = from(handle);
    ARG(c && (wfi == 0 || wfi == 1) && c->wf[wfi].set && c->chan_set && n_pilots >= 1 && M && removed_max >= 0);
    CK(cudaSetDevice(c->device));
    Waveform& w = c->wf[wfi];
    CK(w.Mq.upload(reinterpret_cast<const cplx*>(M), (size_t)n_pilots * c->T * c->N, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    w.mq_P = n_pilots; w.mq_corner = false; w.rsup_zeroed_max = removed_max;
    c->ctas_for_batch = -1;
    return CHEST_OK;
}

int chest_set_estimator_factors(uint64_t handle, int si, int variant, int n_snr, const double* R_inv, double removed_max) {
    Ctx* c = from(handle);
    ARG(c && si >= 0 && si < 3 && (variant == 0 || variant == 1) && R_inv && removed_max >= 0);
    ARG(c->sch[si].set && n_snr == c->S);
    CK(cudaSetDevice(c->device));
    MmseVariant& m = c->sch[si].mm[variant];
    CK(m.rinv.upload(reinterpret_cast<const cplx*>(R_inv), (size_t)n_snr * c->sch[si].P * c->sch[si].P, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    m.rinv_set = true; m.w_zeroed_max = removed_max;
    c->ctas_for_batch = -1;
    return CHEST_OK;
}

int chest_estimator_info(uint64_t handle, int si, int* mode, int* factored, double* removed_r, double* removed_w, float* ms) {
    Ctx* c = from(handle);
    ARG(c && si >= 0 && si < 3);
    if (mode) *mode = c->est_mode;
    if (factored) *factored = c->est_fact[si] ? 1 : 0;
    if (removed_r) *removed_r = c->sch[si].set ? c->wf[c->sch[si].waveform].rsup_zeroed_max : 0.0;
    if (removed_w) *removed_w = std::max(c->sch[si].mm[0].w_zeroed_max, c->sch[si].mm[1].w_zeroed_max);
    if (ms) *ms = c->est_fact_ms;
    return CHEST_OK;
}

int chest_unit_count(uint64_t handle, int* n_units) {
    Ctx* c = from(handle);
    ARG(c && n_units);
    *n_units = c->n_ctas;
    return CHEST_OK;
}

int chest_kernel_times(uint64_t handle, float* ms) {
    Ctx* c = from(handle);
    ARG(c && ms);
    c->kernel_ms[0] = c->hg_ms;
    for (int i = 0; i < 10; ++i) ms[i] = c->kernel_ms[i];
    return CHEST_OK;
}

int chest_work_model(uint64_t handle, int n_iter, double* out) {
    Ctx* c = from(handle);
    ARG(c && out);
    for (int i = 0; i < 8; ++i) out[i] = 0;
    const int S = c->S;
    for (int wfi = 0; wfi < 2; ++wfi) {
        Waveform& w = c->wf[wfi];
        if (!w.set || !w.nsch) continue;
        out[0] += w.flops_d;                                                   // K2
        // perfect CSI (D - diag h) v over D's structural support: 8 flops per complex multiply-add, 4 where v is
        // exactly real (real precoder and constellation)
        for (int q = 0; q < w.nsch; ++q) {
            const Scheme& sq = c->sch[w.sch[q]];
            const bool vr = sq.c_real && c->cst[sq.constellation].real;
            out[2] += (vr ? 4.0 : 8.0) * w.d_struct_pairs * S * n_iter;
        }
        out[3] += w.nsch * (w.flops_mod + 8.0 * c->T * c->N) + w.nsch * S * w.flops_demod;   // TX + demod
        // factored perfect CSI: per column G v, H (.), Q^H (.) over the supports, per iteration
        out[7] += (double)w.nsch * S * n_iter * (w.flops_mod + 8.0 * c->T * c->N + w.flops_demod);
    }
    for (int si = 0; si < 3; ++si) {
        Scheme& s = c->sch[si];
        if (!s.set) continue;
        // estimated CSI: off-diagonal W products of iterations 1..n_iter (D-hat of it-1) + diagonal of 0..n_iter
        double off = 0;
        for (int it = 1; it <= n_iter; ++it) {
            int var_prev = (it - 1 == 0 || (it - 1) <= n_iter / 2) ? 0 : 1;
            off += s.mm[var_prev].nnz_offdiag_pairs * (8.0 * s.P + (s.c_real && c->cst[s.constellation].real ? 4.0 : 8.0));
        }
        out[1] += S * (off + 8.0 * s.K * s.P * (n_iter + 1));
        out[6] += S * off;                                                     // of which in k_ic_main (off-diagonal products)
        out[4] += (double)S * s.mm[0].n_tiles * ((s.P + 3) / 4) * 32 * 16;     // W bytes streamed per IC launch
        out[5] += S * n_iter * (8.0 * s.c_nnz);                                // precoding C z, x2 (est + perfect)
    }
    out[5] *= 2;
    return CHEST_OK;
}

int chest_event_record(uint64_t handle, int slot) {
    Ctx* c = from(handle);
    ARG(c && slot >= 0 && slot < 4);
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->user_ev[slot], c->stream));
    return CHEST_OK;
}
int chest_event_elapsed(uint64_t handle, int a, int b, float* ms) {
    Ctx* c = from(handle);
    ARG(c && a >= 0 && a < 4 && b >= 0 && b < 4 && ms);
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->user_ev[b]));
    CK(cudaEventElapsedTime(ms, c->user_ev[a], c->user_ev[b]));
    return CHEST_OK;
}

int chest_fp64_peak(uint64_t handle, int mode, int iters, double* tflops) {
    Ctx* c = from(handle);
    ARG(c && tflops && iters > 0);
    CK(cudaSetDevice(c->device));
    CK(c->probe.alloc(8));
    const int blocks = c->n_sm * 4, threads = 256;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(a, c->stream));
        if (mode == 0) k_peak_dmma<<<blocks, threads, 0, c->stream>>>(c->probe.p, iters);
        else if (mode == 2) k_peak_mix<<<blocks, threads, 0, c->stream>>>(c->probe.p, iters);
        else k_peak_dfma<<<blocks, threads, 0, c->stream>>>(c->probe.p, iters);
        CK(cudaEventRecord(b, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    c->launches += 2;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    double flops = mode == 0 ? (double)blocks * (threads / 32) * iters * 8.0 * 512.0
                   : mode == 2 ? (double)blocks * (threads / 32) * iters * (4.0 * 512.0 + 8.0 * 64.0)
                               : (double)blocks * threads * iters * 16.0 * 2.0;
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(a); cudaEventDestroy(b);
    return CHEST_OK;
}

}  // extern "C"
