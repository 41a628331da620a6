// kernels.cuh -- the device side of the Monte-Carlo hot path (DS.m:350-565).
//   K1  k_synth_h / k_apply_h      channel synthesis + banded, never-materialised H
//   K2  k_apply_hg, k_gemm_d       D = Q^H (H G) on FP64 tensor cores (DMMA): persistent, support-aware tiles
//   K3  k_gemm<GEMM_DEMOD>, k_estimate   demodulation GEMM over realizations, explicit D-hat
//   K4  k_ic_main, k_ic_light      tensor phase / scalar phases of one interference-cancellation iteration
#pragma once
#include "common.cuh"

#ifndef IC_INLINE
#define IC_INLINE __forceinline__
#endif
#define NC_MAX 16          // columns (realizations x SNR points) one IC CTA carries

// ============================================================================ RNG kernels
// uniforms: out[rep][n_per_rep]
__global__ void k_rng_uniform(double* __restrict__ out, int n_per_rep, int n_rep, int stream,
                              uint64_t seed, int64_t first_rep) {
    int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // two doubles per Philox call
    int rep = blockIdx.y;
    int64_t e = pair * 2;
    if (rep >= n_rep || e >= n_per_rep) return;
    uint64_t r = (uint64_t)(first_rep + rep);
    Philox4 p = philox4x32_10((uint32_t)pair, (uint32_t)r, (uint32_t)stream, (uint32_t)(r >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    double* o = out + (int64_t)rep * n_per_rep;
    o[e] = u53(p.v[0], p.v[1]);
    if (e + 1 < n_per_rep) o[e + 1] = u53(p.v[2], p.v[3]);
}
// bits: out[rep][n_bits] bytes; 128 bits per Philox call
__global__ void k_rng_bits(uint8_t* __restrict__ out, int n_bits, int n_rep, int stream,
                           uint64_t seed, int64_t first_rep) {
    int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int rep = blockIdx.y;
    if (rep >= n_rep || blk * 128 >= n_bits) return;
    uint64_t r = (uint64_t)(first_rep + rep);
    Philox4 p = philox4x32_10((uint32_t)blk, (uint32_t)r, (uint32_t)stream, (uint32_t)(r >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    uint8_t* o = out + (int64_t)rep * n_bits;
    for (int t = 0; t < 128; ++t) {
        int64_t e = blk * 128 + t;
        if (e < n_bits) o[e] = (p.v[t >> 5] >> (t & 31)) & 1u;
    }
}
// indices in [0, order): out[rep][n]
__global__ void k_rng_index(int32_t* __restrict__ out, int n, int n_rep, int order, int stream,
                            uint64_t seed, int64_t first_rep) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    int rep = blockIdx.y;
    if (rep >= n_rep || e >= n) return;
    uint64_t r = (uint64_t)(first_rep + rep);
    Philox4 p = philox4x32_10((uint32_t)(e >> 1), (uint32_t)r, (uint32_t)stream, (uint32_t)(r >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    double u = (e & 1) ? u53(p.v[2], p.v[3]) : u53(p.v[0], p.v[1]);
    int v = (int)floor(u * order);
    out[(int64_t)rep * n + e] = v < order ? v : order - 1;
}
// complex standard normals (Box-Muller): out[rep][snr][n]
__global__ void k_rng_normal(cplx* __restrict__ out, int n, int n_snr, int n_rep,
                             uint64_t seed, int64_t first_rep) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    int snr = blockIdx.y, rep = blockIdx.z;
    if (rep >= n_rep || e >= n) return;
    uint64_t r = (uint64_t)(first_rep + rep);
    Philox4 p = philox4x32_10((uint32_t)e, (uint32_t)r, (uint32_t)(RS_NOISE | (snr << 8)), (uint32_t)(r >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    double u1 = u53(p.v[0], p.v[1]), u2 = u53(p.v[2], p.v[3]);
    double rad = sqrt(-2.0 * log(u1)), s, c;
    sincospi(2.0 * u2, &s, &c);
    out[((int64_t)rep * n_snr + snr) * n + e] = cmake(rad * c, rad * s);
}

// ============================================================================ K1: channel
// h[rep][tap][n] = sqrt(pdp_tap) / sqrt(paths) * sum_p exp(j 2 pi (phase_p + shift_p n dt))
// (FF.m:227-237).  A thread owns SYNTH_SEG consecutive samples: per path one sincos at the first sample, then the phasor
// is advanced by the path's per-sample rotation exp(j 2 pi shift_p dt) (computed once per block, shared memory) -- a complex
// multiplication per sample instead of a sincos.  The recurrence restarts every SYNTH_SEG samples, so its rounding error
// stays below SYNTH_SEG ulp per term (parity with the oracle's exp() at 1e-12 holds with three digits to spare).
#define SYNTH_SEG 16
#define SYNTH_THREADS 64
__global__ void __launch_bounds__(SYNTH_THREADS) k_synth_h(cplx* __restrict__ h, const double* __restrict__ doppler_u,
                          const double* __restrict__ phase_u, const double* __restrict__ tap_amp,
                          int N, int T, int paths, double fD, double dt, int model) {
    extern __shared__ double sm[];
    double* shift = sm;
    double* phase = sm + paths;
    double* rot_c = sm + 2 * paths;
    double* rot_s = sm + 3 * paths;
    int tap = blockIdx.y, rep = blockIdx.z;
    const double* du = doppler_u + (int64_t)rep * T * paths;
    const double* pu = phase_u + (int64_t)rep * T * paths;
    const double PI = 3.14159265358979323846;
    for (int p = threadIdx.x; p < paths; p += blockDim.x) {
        double u = du[tap + T * p];                       // rand([T 1 paths]): tap fastest
        const double sh = (model == 0) ? cos(u * 2.0 * PI) * fD : 2.0 * (u - 0.5) * fD;
        shift[p] = sh;
        phase[p] = pu[tap + T * p];
        double s_, c_;
        sincos((2.0 * PI) * (sh * dt), &s_, &c_);
        rot_c[p] = c_; rot_s[p] = s_;
    }
    __syncthreads();
    const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * SYNTH_SEG;
    if (n0 >= N) return;
    const double t0 = n0 * dt;
    double sr[SYNTH_SEG], si[SYNTH_SEG];
#pragma unroll
    for (int q = 0; q < SYNTH_SEG; ++q) { sr[q] = 0.0; si[q] = 0.0; }
    for (int p = 0; p < paths; ++p) {
        double s_, c_;
        sincos((2.0 * PI) * (phase[p] + shift[p] * t0), &s_, &c_);
        const double rc = rot_c[p], rs = rot_s[p];
#pragma unroll
        for (int q = 0; q < SYNTH_SEG; ++q) {
            sr[q] += c_; si[q] += s_;
            const double cn = fma(c_, rc, -(s_ * rs));
            s_ = fma(s_, rc, c_ * rs);
            c_ = cn;
        }
    }
    const double inv = sqrt((double)paths);
    const double a = tap_amp[tap];
    cplx* out = h + ((int64_t)rep * T + tap) * N + n0;
#pragma unroll
    for (int q = 0; q < SYNTH_SEG; ++q)
        if (n0 + q < N) out[q] = cmake(a * (sr[q] / inv), a * (si[q] / inv));
}

// 'Discrete-Jakes' / 'Discrete-Uniform' (FF.m:203-221): the impulse response of a tap is the inverse DFT of a spectrum with
// 2 n_shift + 1 non-zero bins (Doppler shifts -n_shift .. n_shift times fs/N), i.e. a pruned inverse DFT:
//   h[n] = sum_b coef[b] * gauss[b] * exp(j 2 pi f_b n / N),   coef[b] = sqrt(spectrum_b * pdp_tap / 2)
// gauss: [rep][2 n_shift + 1][T] complex standard normals in the reference's order: rows 0..n_shift are GaussUncorr1
// (shifts 0..n_shift), rows n_shift+1 .. 2 n_shift are GaussUncorr2 (shifts -n_shift .. -1).
__global__ void k_synth_h_discrete(cplx* __restrict__ h, const cplx* __restrict__ gauss, const double* __restrict__ coef,
                                   int N, int T, int n_shift) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x, tap = blockIdx.y, rep = blockIdx.z;
    if (n >= N) return;
    const int nb = 2 * n_shift + 1;
    const cplx* gz = gauss + (int64_t)rep * nb * T;
    cplx acc = cmake(0.0, 0.0);
    for (int b = 0; b < nb; ++b) {
        const int f = b <= n_shift ? b : b - nb;                  // Doppler bin: 0..n_shift, -n_shift..-1
        const int64_t m = (((int64_t)f * n) % N + N) % N;
        double s_, c_;
        sincospi(2.0 * (double)m / (double)N, &s_, &c_);
        const cplx a = gz[b * T + tap];
        const double w = coef[b * T + tap];
        cfma(acc, cmake(a.x * w, a.y * w), cmake(c_, s_));
    }
    h[((int64_t)rep * T + tap) * N + n] = acc;
}
// complex standard normals: out[rep][n_per_rep]
__global__ void k_rng_cnormal(cplx* __restrict__ out, int n_per_rep, int n_rep, int stream, uint64_t seed, int64_t first_rep,
                              double scale = 1.0) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, rep = blockIdx.y;
    if (rep >= n_rep || e >= n_per_rep) return;
    const uint64_t r = (uint64_t)(first_rep + rep);
    Philox4 p = philox4x32_10((uint32_t)e, (uint32_t)r, (uint32_t)stream, (uint32_t)(r >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = u53(p.v[0], p.v[1]), u2 = u53(p.v[2], p.v[3]);
    double rad = sqrt(-2.0 * log(u1)), s_, c_;
    sincospi(2.0 * u2, &s_, &c_);
    out[(int64_t)rep * n_per_rep + e] = cmake(scale * rad * c_, scale * rad * s_);
}

// r[col][n] = sum_tap h[rep(col)][tap][n] * s[col][n - delay_tap]   (banded H, never materialised)
// cols are laid out [group][rep]: rep = col % n_rep.
__global__ void k_apply_h(cplx* __restrict__ r, const cplx* __restrict__ s, const cplx* __restrict__ h,
                          const int* __restrict__ tap_delay, int N, int T, int n_rep, int rep_fixed) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int col = blockIdx.y;
    if (n >= N) return;
    int rep = rep_fixed >= 0 ? rep_fixed : col % n_rep;
    const cplx* sc = s + (int64_t)col * N;
    cplx acc = cmake(0.0, 0.0);
    for (int t = 0; t < T; ++t) {
        int d = tap_delay[t];
        if (n >= d) cfma(acc, h[((int64_t)rep * T + t) * N + n], sc[n - d]);
    }
    r[(int64_t)col * N + n] = acc;
}

// HG[rep][j][n] = sum_tap h[rep][tap][n] * G[n - delay_tap, j] for n inside the k-range of j's column
// tile (zero elsewhere in that range because G is zero outside its support).  K2 then reads H*G as a
// plain k-contiguous operand with cp.async.  The values are stored as the three-multiplication operands of
// K2: plane 1 (complex) = (re, re + im), plane 2 (double) = im - re.  One block handles HG_COLS columns of
// one realization (the channel taps stay in L1 across them).
#define HG_COLS 8
__global__ void k_apply_hg(cplx* __restrict__ HG1, double* __restrict__ HG2, const cplx* __restrict__ G,
                           const cplx* __restrict__ h, const int* __restrict__ tap_delay,
                           const int* __restrict__ nt_klo, const int* __restrict__ nt_khi, int N, int Np, int K, int T,
                           int rep0, int tile) {
    const int rep = blockIdx.y + rep0;
    const cplx* hr = h + (int64_t)rep * T * N;
    for (int jj = 0; jj < HG_COLS; ++jj) {
        const int j = blockIdx.x * HG_COLS + jj;
        if (j >= K) break;
        // K2 copies the planes in aligned pairs: cover the range rounded out to even bounds (exact zeros there)
        const int lo = nt_klo[j / tile] & ~1, hi = min(N, (nt_khi[j / tile] + 1) & ~1);
        const cplx* gc = G + (int64_t)N * j;
        const int64_t o = ((int64_t)rep * K + j) * Np;
        for (int n = lo + threadIdx.x; n < hi; n += blockDim.x) {
            cplx acc = cmake(0.0, 0.0);
            for (int t = 0; t < T; ++t) {
                int d = tap_delay[t];
                if (n >= d) cfma(acc, hr[(int64_t)t * N + n], gc[n - d]);
            }
            HG1[o + n] = cmake(acc.x, acc.x + acc.y);
            HG2[o + n] = acc.y - acc.x;
        }
    }
}

// r[col][n] = sum_tap h[rep_of_col[col]][tap][n] * s[col][n - delay_tap]: the banded H applied to a list of columns
__global__ void k_apply_h_cols(cplx* __restrict__ r, const cplx* __restrict__ s, const cplx* __restrict__ h,
                               const int* __restrict__ tap_delay, const int* __restrict__ rep_of_col, int N, int T) {
    const int n = blockIdx.y * blockDim.x + threadIdx.x, col = blockIdx.x;      // columns on grid.x: up to 2^31 - 1 of them
    if (n >= N) return;
    const int rep = rep_of_col[col];
    const cplx* sc = s + (int64_t)col * N;
    cplx acc = cmake(0.0, 0.0);
    for (int t = 0; t < T; ++t) {
        const int d = tap_delay[t];
        if (n >= d) cfma(acc, h[((int64_t)rep * T + t) * N + n], sc[n - d]);
    }
    r[(int64_t)col * N + n] = acc;
}

// the same, r written as the three-multiplication operand planes of the next GEMM: r1 = (re, re + im), r2 = im - re; row stride Np
__global__ void k_apply_h_cols_planes(cplx* __restrict__ r1, double* __restrict__ r2, const cplx* __restrict__ s, const cplx* __restrict__ h,
                                      const int* __restrict__ tap_delay, const int* __restrict__ rep_of_col, int N, int Np, int T) {
    const int n = blockIdx.y * blockDim.x + threadIdx.x, col = blockIdx.x;      // columns on grid.x: up to 2^31 - 1 of them
    if (n >= N) return;
    const int rep = rep_of_col[col];
    const cplx* sc = s + (int64_t)col * N;
    cplx acc = cmake(0.0, 0.0);
    for (int t = 0; t < T; ++t) {
        const int d = tap_delay[t];
        if (n >= d) cfma(acc, h[((int64_t)rep * T + t) * N + n], sc[n - d]);
    }
    r1[(int64_t)col * Np + n] = cmake(acc.x, acc.x + acc.y);
    r2[(int64_t)col * Np + n] = acc.y - acc.x;
}

// ============================================================================ constellations
struct ConstDev {
    int order, nbits, n_axis;           // n_axis = order (PAM) or sqrt(order) (QAM)
    int is_qam;
    const cplx* symbol;                 // SymbolMapping[word]
    const cplx* pilot;                  // SymbolMapping[word] / |.|
    const double* level;                // axis levels, ascending (n_axis)
    double inv_step;                    // (n_axis - 1) / (level[n_axis-1] - level[0])
    const int* word_of_grid;            // PAM: [t] ; QAM: [ti * n_axis + tq] -> word
};

// nearest axis level, decided by the same |x - level| comparison the reference's argmin uses
__device__ __forceinline__ int nearest_level(const double* lev, int n, double inv_step, double x) {
    // first guess from the uniform grid (reciprocal multiply: the guess only has to be within one level,
    // the decision itself is made by the |x - level| comparisons below)
    int t = (int)floor((x - lev[0]) * inv_step + 0.5);
    t = t < 0 ? 0 : (t > n - 1 ? n - 1 : t);
    double best = fabs(x - lev[t]);
    if (t > 0 && fabs(x - lev[t - 1]) < best) { best = fabs(x - lev[t - 1]); t = t - 1; }
    else if (t < n - 1 && fabs(x - lev[t + 1]) < best) { t = t + 1; }
    return t;
}
__device__ __forceinline__ int demap_word(const ConstDev& cd, cplx x) {
    if (!cd.is_qam) return cd.word_of_grid[nearest_level(cd.level, cd.n_axis, cd.inv_step, x.x)];
    int ti = nearest_level(cd.level, cd.n_axis, cd.inv_step, x.x);
    int tq = nearest_level(cd.level, cd.n_axis, cd.inv_step, x.y);
    return cd.word_of_grid[ti * cd.n_axis + tq];
}

// ============================================================================ scheme descriptors
struct WTiles {                    // one (variant, snr) MMSE matrix in diagonal-tile form
    const cplx* frag;              // [tile][P4][32 lanes] : lane = 4*row + p%4  -> W[i, i+delta, p]
    const cplx* diag;              // [K][P]  W[i,i,p]
};
struct SchemeDev {
    int waveform, K, K_in, P, P4, n_data, nbits, detect_mode, constellation, n_bits_total;
    int v_real;                    // precoder and constellation are exactly real: v = C z has a zero imaginary part
    double sqrt_kappa, dpr, sqrt_dpr, inv_sqrt_dpr, inv_dpr;
    const int* c_rowptr; const int* c_col; const cplx* c_val;      // precoder C, CSR (K rows)
    const int* ct_colptr; const int* ct_row; const cplx* ct_val;   // precoder C, CSC (K_in cols)
    const int* pilot_pos; const int* data_pos; const uint32_t* edge_mask;   // per data symbol bit mask
    const int* pos2data;           // [K] data-symbol index read off position i (select schemes), -1 otherwise
    // precoder rows with a single entry (ELL-1): column (-1: empty row, -2: long row) and value; long rows listed
    const int* row_col0; const cplx* row_val0; const int* long_rows; int n_long_rows;
    // long rows as DMMA tiles: 8 long rows per tile, the union of their columns packed four per k-step;
    // lr_ptr [n_lr_tiles+1] -> steps, lr_kcol [step][4] column of z, lr_frag [step][32 lanes] A fragments
    int n_lr_tiles; const int* lr_ptr; const int* lr_kcol; const cplx* lr_frag;
    // k_ic_post: per-row descriptor {d decided directly at this row (-1: none), col0, NoEdge mask of d, flags (bit 0:
    // the equalised symbol of this row is needed by the de-spreading gather)}, the data symbols the gather pass decides,
    // and whether v of every single-entry row can be written in the same pass as the decision
    const int4* rowinfo; const int* multi_d; int n_multi; int fuse_ok;
    const uint32_t* txw_t;         // transmitted words, [rep / 16][n_data][rep % 16]
    const cplx* wdiag_frag[2];     // [snr][rt][pq][32 lanes]: W[i,i,p] in DMMA A-fragment order (phase D)
    // MMSE matrices: tile lists per variant, fragments per (variant, snr)
    const int* tile_ptr[2];        // [RT+1]
    const int* tile_delta[2];      // [n_tiles]
    const WTiles* w[2];            // [n_snr]
    // per-batch state
    cplx* xP;                      // [rep][P]      unit-modulus pilots
    uint32_t* txword;              // [rep][n_data] transmitted bit words
    uint32_t* txw_t_w;             // the same, [rep / 16][n_data][rep % 16] (written by k_tx_symbols)
    cplx* x;                       // [rep][K]      precoded symbols
    cplx* y;                       // [snr][rep][K]
    cplx* hP;                      // [snr][rep][P]
    cplx* hdiag;                   // [snr][rep][K]
    cplx* xD[2];                   // [csi][snr][rep][n_data]
};

// ============================================================================ TX symbols
// bits -> data symbols, pilot indices -> unit-modulus pilots, x = C [xP; xD]  (DS.m:355-373)
__global__ void k_tx_symbols(SchemeDev sd, ConstDev cd, const uint8_t* __restrict__ bits,
                             const int32_t* __restrict__ pilot_idx, int n_rep) {
    extern __shared__ cplx zs[];
    int rep = blockIdx.x;
    if (rep >= n_rep) return;
    const uint8_t* b = bits + (int64_t)rep * sd.n_bits_total;
    for (int k = threadIdx.x; k < sd.K_in; k += blockDim.x) {
        cplx z;
        if (k < sd.P) {
            z = cd.pilot[pilot_idx[(int64_t)rep * sd.P + k]];
            sd.xP[(int64_t)rep * sd.P + k] = z;
        } else {
            int d = k - sd.P;
            uint32_t w = 0;
            for (int t = 0; t < sd.nbits; ++t) w |= (uint32_t)(b[d * sd.nbits + t] & 1) << t;
            sd.txword[(int64_t)rep * sd.n_data + d] = w;
            sd.txw_t_w[((int64_t)(rep >> 4) * sd.n_data + d) * 16 + (rep & 15)] = w;
            z = cd.symbol[w];
        }
        zs[k] = z;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < sd.K; i += blockDim.x) {
        cplx acc = cmake(0.0, 0.0);
        for (int e = sd.c_rowptr[i]; e < sd.c_rowptr[i + 1]; ++e) cfma(acc, sd.c_val[e], zs[sd.c_col[e]]);
        sd.x[(int64_t)rep * sd.K + i] = acc;
    }
}

// ============================================================================ shared-A complex GEMM
// C[m, col] = sum_k opA[m,k] * B[k, col]  with A shared by every column (and every realization):
//   A is given "k-contiguous": At[k + lda*m]; conj flag folds Q^H.
//   B[k, col] comes from a mode-specific loader; C is written as out[col*ldc + m].
// 64x64 CTA tile, 8 warps (2x4), warp tile 32x16 = 4x2 DMMA tiles, KT = 32.
enum { GEMM_PLAIN = 0, GEMM_DEMOD = 1 };
struct GemmParams {
    int M, Kc, n_cols, lda, ldc, conj_a;
    const cplx* At;
    // GEMM_DEMOD (A = Q^H): three-multiplication planes of the conjugated operand, same indexing as At:
    // At1 = (re, re - im), At2 = im
    const cplx* At1; const double* At2;
    const int* mt_klo; const int* mt_khi;      // per CTA row tile: k support range
    const int* m8_klo; const int* m8_khi;      // per 8 rows: k support range (warp-level clipping; may be null)
    cplx* out;                                  // [col][ldc]
    // PLAIN: B[k,col] = bsrc[col*ldb + k], or with a column table bsrc[b_off[col] + k*b_kstride]
    // (the unit-interleaved v of the IC scratch: k stride 16)
    const cplx* bsrc; int ldb;
    const int64_t* b_off; int b_kstride;
    // PLAIN, cancellation epilogue (factored perfect-CSI pass): instead of out[col][m] = U the kernel writes
    // e_out[e_off[col] + m*16] = e_y[e_yoff[col] + m] - U + e_h[e_rep[col]*M + m] * bsrc2[e_off2[col] + m*16]
    cplx* e_out; const int64_t* e_off; const cplx* e_y; const int64_t* e_yoff; const cplx* e_h; const int* e_rep;
    const cplx* e_v; const int64_t* e_voff;
    // DEMOD: col = (g*n_snr + snr)*n_rep + rep ; B = r0[(g*n_rep+rep)*N + k] + sqrt(pn[snr]/2)*noise[(rep*n_snr+snr)*N + k]
    const cplx* r0; const cplx* noise; const double* noise_scale; int n_snr, n_rep;
    int N;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;                       // src-size 0 -> the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Tiles live in shared memory as interleaved complex [row][k] with a row stride of KT+4 elements, so a
// lane's (re, im) fragment pair is one conflict-free LDS.128.  Two stages: while the tensor pipe works
// on stage s, cp.async (A, and B when it is a plain copy) and the computed B values (noise add) fill
// stage s^1; one barrier per k-tile.
// Geometry: WM x WN warps, each owning TMW x 2 DMMA tiles -> CTA tile (8*TMW*WM) x (16*WN).
//   <2,4,4>: 64 x 64, 8 warps        <2,3,3>: 48 x 48, 6 warps, 2 CTAs/SM
// The host picks, per waveform, the geometry whose tiles waste the fewest flops on padding of D = Q^H (H G)
// (48 divides the 720 / 336 symbols of the reference's grids and follows the 24-subcarrier support structure
// more closely); K2 itself runs in k_gemm_d below, this kernel serves s = G x (PLAIN) and the demodulation.
// Complex products use the three-multiplication form (CHEST_3M, common.cuh): per 8x8 tile the partial sums
// c1 = (ar+ai) br, cr = -ai (br+bi), ci = ar (bi-br); the epilogue forms re = c1 + cr, im = c1 + ci.
// GEMM_DEMOD (PRE3) reads the operand sums ready-made: A from the planes the host builds once per waveform,
// B computed while adding the noise, so the inner loop is LDS + DMMA only; a second, double-valued smem plane per operand holds the third value (row stride 20
// doubles: conflict-free 8-byte fragment loads).  GEMM_PLAIN forms the sums in registers.
template <int MODE, int WM, int WN, int TMW>
__global__ void __launch_bounds__(32 * WM * WN, CHEST_3M ? ((WM * WN == 8) ? 1 : 2) : ((WM * WN == 8) ? 2 : 3))
k_gemm(GemmParams p) {
    constexpr int TM = 8 * TMW * WM, TN = 16 * WN, KT = 16, LDS = KT + 4, NTHR = 32 * WM * WN;
    constexpr bool PRE3 = CHEST_3M && MODE != GEMM_PLAIN;
    extern __shared__ __align__(128) double smem[];
    cplx (*As)[TM][LDS] = reinterpret_cast<cplx (*)[TM][LDS]>(smem);
    cplx (*Bs)[TN][LDS] = reinterpret_cast<cplx (*)[TN][LDS]>(smem + 2 * 2 * TM * LDS);
    double (*As2)[TM][LDS] = reinterpret_cast<double (*)[TM][LDS]>(smem + 2 * 2 * (TM + TN) * LDS);
    double (*Bs2)[TN][LDS] = reinterpret_cast<double (*)[TN][LDS]>(smem + 2 * 2 * (TM + TN) * LDS + 2 * TM * LDS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t4 = lane & 3;
    const int mt = blockIdx.x, nt = blockIdx.y;
    const int m0 = mt * TM, n0 = nt * TN;
    int klo = p.mt_klo ? p.mt_klo[mt] : 0, khi = p.mt_khi ? p.mt_khi[mt] : p.Kc;
    if (PRE3) klo &= ~1;                                  // the double planes are copied in aligned pairs
    cplx* out = p.out;
    const bool conj_a = p.conj_a != 0;

    double cr[TMW][2][2], ci[TMW][2][2], c1[CHEST_3M ? TMW : 1][2][2];
#pragma unroll
    for (int a = 0; a < TMW; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            cr[a][b][0] = cr[a][b][1] = ci[a][b][0] = ci[a][b][1] = 0.0;
            if (CHEST_3M) c1[a][b][0] = c1[a][b][1] = 0.0;
        }

    auto fill = [&](int stage, int k0) {
#pragma unroll
        for (int e = 0; e < (TM * KT) / NTHR; ++e) {         // A tile: TM rows x 16 k
            int idx = tid + e * NTHR, kk = idx & (KT - 1), r = idx / KT;
            int gk = k0 + kk, gm = m0 + r;
            bool ok = gm < p.M && gk < khi;
            cp_async16(&As[stage][r][kk], (PRE3 ? p.At1 : p.At) + (ok ? (int64_t)gk + (int64_t)p.lda * gm : 0), ok);
        }
        if (PRE3) {
#pragma unroll
            for (int e = 0; e < (TM * KT / 2) / NTHR; ++e) { // third A value: pairs of doubles
                int idx = tid + e * NTHR, kk = (idx & (KT / 2 - 1)) * 2, r = idx / (KT / 2);
                int gk = k0 + kk, gm = m0 + r;
                bool ok = gm < p.M && gk < khi;
                cp_async16(&As2[stage][r][kk], p.At2 + (ok ? (int64_t)gk + (int64_t)p.lda * gm : 0), ok);
            }
        }
#pragma unroll
        for (int e = 0; e < (TN * KT) / NTHR; ++e) {         // B tile: TN cols x 16 k
            int idx = tid + e * NTHR, kk = idx & (KT - 1), c = idx / KT;
            int gk = k0 + kk, col = n0 + c;
            bool ok = col < p.n_cols && gk < khi;
            if (MODE == GEMM_PLAIN) {
                const int64_t o = !ok ? 0 : (p.b_off ? p.b_off[col] + (int64_t)gk * p.b_kstride : (int64_t)col * p.ldb + gk);
                cp_async16(&Bs[stage][c][kk], p.bsrc + o, ok);
            } else {
                cplx v = cmake(0.0, 0.0);
                if (ok) {
                    int r_ = col % p.n_rep, q = col / p.n_rep, snr = q % p.n_snr, grp = q / p.n_snr;
                    cplx a = p.r0[((int64_t)grp * p.n_rep + r_) * p.N + gk];
                    cplx nz = p.noise[((int64_t)r_ * p.n_snr + snr) * p.N + gk];
                    double sc = p.noise_scale[snr];
                    v = cmake(a.x + sc * nz.x, a.y + sc * nz.y);
                }
                if (PRE3) { Bs[stage][c][kk] = cmake(v.x, v.x + v.y); Bs2[stage][c][kk] = v.y - v.x; }
                else Bs[stage][c][kk] = v;
            }
        }
    };

    // warp-level clipping: a warp skips the k-steps outside the support of its own rows / columns (the CTA range
    // is the union over 2 x 3 warp tiles; for the FBMC grid a warp's 24 rows are one time position)
    int wlo = klo, whi = khi;
    if (p.m8_klo) {
        int lo = 0x7fffffff, hi = 0;
#pragma unroll
        for (int x = 0; x < TMW; ++x) {
            const int r = m0 + wm * 8 * TMW + x * 8;
            if (r < p.M) { lo = min(lo, p.m8_klo[r >> 3]); hi = max(hi, p.m8_khi[r >> 3]); }
        }
        wlo = max(wlo, lo); whi = min(whi, hi);
    }
    const int nk = (khi - klo + KT - 1) / KT;
    if (nk > 0) {
        fill(0, klo);
        cp_async_wait_all();
        __syncthreads();
    }
    for (int kt = 0; kt < nk; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < nk) fill(st ^ 1, klo + (kt + 1) * KT);
        const int kbase = klo + kt * KT;
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            if (kbase + kk + 4 <= wlo || kbase + kk >= whi) continue;     // warp-uniform
            cplx a[TMW], b[2];
#pragma unroll
            for (int x = 0; x < TMW; ++x) a[x] = As[st][wm * 8 * TMW + x * 8 + g][kk + t4];
#pragma unroll
            for (int y = 0; y < 2; ++y) b[y] = Bs[st][wn * 16 + y * 8 + g][kk + t4];
            if (PRE3) {
                double a2[TMW], b2[2];
#pragma unroll
                for (int x = 0; x < TMW; ++x) a2[x] = As2[st][wm * 8 * TMW + x * 8 + g][kk + t4];
#pragma unroll
                for (int y = 0; y < 2; ++y) b2[y] = Bs2[st][wn * 16 + y * 8 + g][kk + t4];
#pragma unroll
                for (int x = 0; x < TMW; ++x)
#pragma unroll
                    for (int y = 0; y < 2; ++y) {
                        dmma884(c1[CHEST_3M ? x : 0][y][0], c1[CHEST_3M ? x : 0][y][1], a[x].y, b[y].x);   // (ar+ai') br
                        dmma884(cr[x][y][0], cr[x][y][1], a2[x], b[y].y);                                  // -ai' (br+bi)
                        dmma884(ci[x][y][0], ci[x][y][1], a[x].x, b2[y]);                                  // ar (bi-br)
                    }
            } else if (CHEST_3M) {
                double bs[2], bd[2];
#pragma unroll
                for (int y = 0; y < 2; ++y) { bs[y] = b[y].x + b[y].y; bd[y] = b[y].y - b[y].x; }
#pragma unroll
                for (int x = 0; x < TMW; ++x) {
                    const double ai = conj_a ? dneg(a[x].y) : a[x].y;
                    const double nai = conj_a ? a[x].y : dneg(a[x].y);
                    const double as = a[x].x + ai;
#pragma unroll
                    for (int y = 0; y < 2; ++y) {
                        dmma884(c1[CHEST_3M ? x : 0][y][0], c1[CHEST_3M ? x : 0][y][1], as, b[y].x);
                        dmma884(cr[x][y][0], cr[x][y][1], nai, bs[y]);
                        dmma884(ci[x][y][0], ci[x][y][1], a[x].x, bd[y]);
                    }
                }
            } else {
#pragma unroll
                for (int x = 0; x < TMW; ++x) {
                    const double ai = conj_a ? dneg(a[x].y) : a[x].y;
                    const double nai = conj_a ? a[x].y : dneg(a[x].y);
#pragma unroll
                    for (int y = 0; y < 2; ++y) {
                        dmma884(cr[x][y][0], cr[x][y][1], a[x].x, b[y].x);
                        dmma884(cr[x][y][0], cr[x][y][1], nai, b[y].y);
                        dmma884(ci[x][y][0], ci[x][y][1], a[x].x, b[y].y);
                        dmma884(ci[x][y][0], ci[x][y][1], ai, b[y].x);
                    }
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
    }
    // ---- epilogue: C[g][2*t4 + e] of each 8x8 tile
#pragma unroll
    for (int a = 0; a < TMW; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int m = m0 + wm * 8 * TMW + a * 8 + g, col = n0 + wn * 16 + b * 8 + 2 * t4 + e;
                if (m < p.M && col < p.n_cols) {
                    cplx v = CHEST_3M ? cmake(c1[CHEST_3M ? a : 0][b][e] + cr[a][b][e], c1[CHEST_3M ? a : 0][b][e] + ci[a][b][e])
                                      : cmake(cr[a][b][e], ci[a][b][e]);
                    if (MODE == GEMM_PLAIN && p.e_out) {       // y_ic = y - U + h v  (DS.m:541-543 with U = Q^H H G v)
                        const cplx yv = p.e_y[p.e_yoff[col] + m], hv = p.e_h[(int64_t)p.e_rep[col] * p.M + m];
                        const cplx vv = p.e_v[p.e_voff[col] + (int64_t)m * 16];
                        const cplx hvv = cmul(hv, vv);
                        p.e_out[p.e_off[col] + (int64_t)m * 16] = cmake(yv.x - v.x + hvv.x, yv.y - v.y + hvv.y);
                    } else
                    out[(int64_t)col * p.ldc + m] = v;
                }
            }
}

// ============================================================================ K2: D = Q^H (H G), persistent
// The transmission-matrix GEMM as a persistent kernel: every CTA walks a list of (realization, row tile,
// column tile) work items -- only the tile pairs whose supports overlap (the others are structural zeros of D
// and were cleared once at allocation) -- and streams their k-tiles through ONE shared-memory ring that runs
// across tile boundaries: the first operands of the next tile are in flight while the current one finishes, so
// there is no per-tile prologue.  Stages are handed over with mbarriers instead of block barriers:
//   full[s]   every thread's cp.async group for stage s has landed (cp.async.mbarrier.arrive.noinc)
//   empty[s]  every warp has finished reading stage s (one arrive per warp)
// so a warp only ever waits for the data it needs, not for the slowest warp of the previous k-tile.  Operands
// are the three-multiplication planes (see k_gemm): the inner loop is LDS + DMMA.
struct GemmDParams {
    int M, n_cols, lda, ldb, n_rep, rep0, n_pairs;
    const int2* pairs;                           // (row tile, column tile) with overlapping supports
    const cplx* At1; const double* At2;          // Q^H planes [m][lda]
    const cplx* b1; const double* b2;            // H*G planes [rep][col][ldb]
    const int* mt_klo; const int* mt_khi; const int* nt_klo; const int* nt_khi;
    const int* m8_klo; const int* m8_khi; const int* n8_klo; const int* n8_khi;
    cplx* out;                                   // D, row-tile-major [rep][M/8][n_cols][8]
    cplx* hdiag;                                 // [rep][M]
};
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared.b64 st, [%0]; }" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(uint64_t* bar) {   // arrives when this thread's prior cp.async ops are done
    asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n .reg .pred p;\nWAIT_%=:\n mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\nDONE_%=:\n}"
                 :: "r"(a), "r"(parity) : "memory");
}

#ifndef GEMMD_STAGES
#define GEMMD_STAGES 4
#endif
#ifndef GEMMD_KT
#define GEMMD_KT 8             // k-tile per stage (8 or 16)
#endif
#ifndef GEMMD_DEPTH
#define GEMMD_DEPTH (GEMMD_STAGES > 2 ? GEMMD_STAGES - 2 : 1)
#endif
template <int WM, int WN, int TMW>
__global__ void __launch_bounds__(32 * WM * WN, (WM * WN == 8) ? 1 : 2) k_gemm_d(GemmDParams p) {
    constexpr int TM = 8 * TMW * WM, TN = 16 * WN, KT = GEMMD_KT, LDS = KT + 4, NTHR = 32 * WM * WN, NS = GEMMD_STAGES;
    extern __shared__ __align__(128) double smem[];
    cplx (*As)[TM][LDS] = reinterpret_cast<cplx (*)[TM][LDS]>(smem);
    cplx (*Bs)[TN][LDS] = reinterpret_cast<cplx (*)[TN][LDS]>(smem + NS * 2 * TM * LDS);
    double (*As2)[TM][LDS] = reinterpret_cast<double (*)[TM][LDS]>(smem + NS * 2 * (TM + TN) * LDS);
    double (*Bs2)[TN][LDS] = reinterpret_cast<double (*)[TN][LDS]>(smem + NS * 2 * (TM + TN) * LDS + NS * TM * LDS);
    __shared__ uint64_t full[NS], empty[NS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t4 = lane & 3;
    if (tid == 0) {
        for (int s_ = 0; s_ < NS; ++s_) { mbar_init(&full[s_], NTHR); mbar_init(&empty[s_], WM * WN); }
    }
    __syncthreads();
    const int total = p.n_pairs * p.n_rep;                     // work items (host checks the product fits an int)
    const int RT8 = ((p.M + 7) / 8) * 8;

    // ---- consumer cursor: the item whose k-tiles the warps multiply
    struct Cursor { int t, kt, nk, klo, khi, m0, n0, rep; };
    auto load_item = [&](Cursor& c) {
        if (c.t >= total) { c.nk = 0; return; }
        c.rep = c.t / p.n_pairs;
        const int2 mn = p.pairs[c.t - c.rep * p.n_pairs];
        c.rep += p.rep0;
        c.klo = max(p.mt_klo[mn.x], p.nt_klo[mn.y]) & ~1;      // the double planes are copied in aligned pairs
        c.khi = min(p.mt_khi[mn.x], p.nt_khi[mn.y]);
        c.nk = (c.khi - c.klo + KT - 1) / KT;
        c.m0 = mn.x * TM; c.n0 = mn.y * TN; c.kt = 0;
    };
    Cursor cc;
    cc.t = blockIdx.x;
    load_item(cc);

    // ---- producer: every thread copies fixed (row, k) slots of each stage; per item it keeps four source
    //      pointers (advanced by KT per job) and the row-validity flags, so a job costs a handful of instructions
    constexpr int EA = (TM * KT) / NTHR, EB = (TN * KT) / NTHR, EA2 = (TM * KT / 2) / NTHR, EB2 = (TN * KT / 2) / NTHR;
    static_assert(EA * NTHR == TM * KT && EB * NTHR == TN * KT && EA2 * NTHR == TM * KT / 2 && EB2 * NTHR == TN * KT / 2, "tile / thread mismatch");
    constexpr int RSTEP = NTHR / KT, RSTEP2 = NTHR / (KT / 2);         // row distance between a thread's slots
    const int kk1 = tid & (KT - 1), r1 = tid / KT, kk2 = (tid & (KT / 2 - 1)) * 2, r2 = tid / (KT / 2);
    const unsigned sA1 = (unsigned)__cvta_generic_to_shared(&As[0][r1][kk1]), sB1 = (unsigned)__cvta_generic_to_shared(&Bs[0][r1][kk1]);
    const unsigned sA2 = (unsigned)__cvta_generic_to_shared(&As2[0][r2][kk2]), sB2 = (unsigned)__cvta_generic_to_shared(&Bs2[0][r2][kk2]);
    constexpr unsigned STG_A1 = TM * LDS * 16, STG_B1 = TN * LDS * 16, STG_A2 = TM * LDS * 8, STG_B2 = TN * LDS * 8;
    int pt = blockIdx.x, pkt = 0, pnk = 0, prem = 0;           // producer item, k-tile, k-tiles of the item, valid k left
    const cplx* ga1 = nullptr; const cplx* gb1 = nullptr; const double* ga2 = nullptr; const double* gb2 = nullptr;
    unsigned okA1 = 0, okB1 = 0, okA2 = 0, okB2 = 0;           // bit e: slot e of this thread is inside the matrix
    auto producer_item = [&]() {
        if (pt >= total) { pnk = 0; return; }
        int rep = pt / p.n_pairs;
        const int2 mn = p.pairs[pt - rep * p.n_pairs];
        rep += p.rep0;
        const int klo = max(p.mt_klo[mn.x], p.nt_klo[mn.y]) & ~1, khi = min(p.mt_khi[mn.x], p.nt_khi[mn.y]);
        pnk = (khi - klo + KT - 1) / KT; pkt = 0; prem = khi - klo;
        const int m0 = mn.x * TM, n0 = mn.y * TN;
        ga1 = p.At1 + (int64_t)(m0 + r1) * p.lda + klo + kk1;
        ga2 = p.At2 + (int64_t)(m0 + r2) * p.lda + klo + kk2;
        gb1 = p.b1 + ((int64_t)rep * p.n_cols + n0 + r1) * p.ldb + klo + kk1;
        gb2 = p.b2 + ((int64_t)rep * p.n_cols + n0 + r2) * p.ldb + klo + kk2;
        okA1 = okB1 = okA2 = okB2 = 0;
#pragma unroll
        for (int e = 0; e < EA; ++e) okA1 |= (unsigned)(m0 + r1 + e * RSTEP < p.M) << e;
#pragma unroll
        for (int e = 0; e < EB; ++e) okB1 |= (unsigned)(n0 + r1 + e * RSTEP < p.n_cols) << e;
#pragma unroll
        for (int e = 0; e < EA2; ++e) okA2 |= (unsigned)(m0 + r2 + e * RSTEP2 < p.M) << e;
#pragma unroll
        for (int e = 0; e < EB2; ++e) okB2 |= (unsigned)(n0 + r2 + e * RSTEP2 < p.n_cols) << e;
    };
    auto cp16 = [](unsigned dst, const void* src, bool ok) {   // !ok: no read, the 16 bytes are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(src), "r"(ok ? 16 : 0));
    };
    unsigned pj = 0, cj = 0;                                   // k-tile jobs produced / consumed by this thread
    auto produce = [&]() {
        if (pnk == 0) return;
        const unsigned st = pj % NS;
        if (pj >= NS) mbar_wait(&empty[st], ((pj / NS) - 1) & 1);     // every warp is done with the stage's previous job
        const bool k1 = kk1 < prem, k2 = kk2 < prem;
        const int64_t lda = p.lda, ldb = p.ldb;
#pragma unroll
        for (int e = 0; e < EA; ++e) { const bool ok = k1 && ((okA1 >> e) & 1); cp16(sA1 + st * STG_A1 + e * RSTEP * LDS * 16, ok ? ga1 + e * RSTEP * lda : p.At1, ok); }
#pragma unroll
        for (int e = 0; e < EA2; ++e) { const bool ok = k2 && ((okA2 >> e) & 1); cp16(sA2 + st * STG_A2 + e * RSTEP2 * LDS * 8, ok ? ga2 + e * RSTEP2 * lda : p.At2, ok); }
#pragma unroll
        for (int e = 0; e < EB; ++e) { const bool ok = k1 && ((okB1 >> e) & 1); cp16(sB1 + st * STG_B1 + e * RSTEP * LDS * 16, ok ? gb1 + e * RSTEP * ldb : p.b1, ok); }
#pragma unroll
        for (int e = 0; e < EB2; ++e) { const bool ok = k2 && ((okB2 >> e) & 1); cp16(sB2 + st * STG_B2 + e * RSTEP2 * LDS * 8, ok ? gb2 + e * RSTEP2 * ldb : p.b2, ok); }
        mbar_arrive_cp_async(&full[st]);
        ++pj;
        ga1 += KT; ga2 += KT; gb1 += KT; gb2 += KT; prem -= KT;
        if (++pkt == pnk) { pt += gridDim.x; producer_item(); }
    };
    producer_item();

    double cr[TMW][2][2], ci[TMW][2][2], c1[TMW][2][2];
    auto clear_acc = [&]() {
#pragma unroll
        for (int a = 0; a < TMW; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) { cr[a][b][0] = cr[a][b][1] = ci[a][b][0] = ci[a][b][1] = c1[a][b][0] = c1[a][b][1] = 0.0; }
    };
    int wlo = 0, whi = 0;
    auto warp_range = [&]() {                                  // k-steps outside the support of this warp's rows / columns are skipped
        wlo = cc.klo; whi = cc.khi;
        int lo = 0x7fffffff, hi = 0;
#pragma unroll
        for (int x = 0; x < TMW; ++x) {
            const int r = cc.m0 + wm * 8 * TMW + x * 8;
            if (r < p.M) { lo = min(lo, p.m8_klo[r >> 3]); hi = max(hi, p.m8_khi[r >> 3]); }
        }
        wlo = max(wlo, lo); whi = min(whi, hi);
        lo = 0x7fffffff; hi = 0;
#pragma unroll
        for (int y = 0; y < 2; ++y) {
            const int c = cc.n0 + wn * 16 + y * 8;
            if (c < p.n_cols) { lo = min(lo, p.n8_klo[c >> 3]); hi = max(hi, p.n8_khi[c >> 3]); }
        }
        wlo = max(wlo, lo); whi = min(whi, hi);
    };
    clear_acc();
    if (cc.t < total) warp_range();
#pragma unroll 1
    // the producer runs GEMMD_DEPTH jobs ahead; with DEPTH < NS - 1 the stage it refills was freed NS - DEPTH jobs ago,
    // so a warp is not held up by warps that are still on the previous job (DEPTH = NS - 1 is a barrier per job)
    for (int i = 0; i < GEMMD_DEPTH; ++i) produce();
#pragma unroll 1
    while (cc.t < total) {
        produce();
        const int st = cj % NS;
        mbar_wait(&full[st], (cj / NS) & 1);
        const int kbase = cc.klo + cc.kt * KT;
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            if (kbase + kk + 4 <= wlo || kbase + kk >= whi) continue;     // warp-uniform
            cplx a[TMW], b[2];
            double a2[TMW], b2[2];
#pragma unroll
            for (int x = 0; x < TMW; ++x) { a[x] = As[st][wm * 8 * TMW + x * 8 + g][kk + t4]; a2[x] = As2[st][wm * 8 * TMW + x * 8 + g][kk + t4]; }
#pragma unroll
            for (int y = 0; y < 2; ++y) { b[y] = Bs[st][wn * 16 + y * 8 + g][kk + t4]; b2[y] = Bs2[st][wn * 16 + y * 8 + g][kk + t4]; }
#pragma unroll
            for (int x = 0; x < TMW; ++x)
#pragma unroll
                for (int y = 0; y < 2; ++y) {
                    dmma884(c1[x][y][0], c1[x][y][1], a[x].y, b[y].x);     // (ar+ai') br
                    dmma884(cr[x][y][0], cr[x][y][1], a2[x], b[y].y);      // -ai' (br+bi)
                    dmma884(ci[x][y][0], ci[x][y][1], a[x].x, b2[y]);      // ar (bi-br)
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        ++cj;
        if (++cc.kt == cc.nk) {
            // ---- epilogue of this tile: D row-tile-major, diagonal to hdiag
            cplx* out = p.out + (int64_t)cc.rep * RT8 * p.n_cols;
#pragma unroll
            for (int a = 0; a < TMW; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int m = cc.m0 + wm * 8 * TMW + a * 8 + g, col = cc.n0 + wn * 16 + b * 8 + 2 * t4 + e;
                        if (m < p.M && col < p.n_cols) {
                            const cplx v = cmake(c1[a][b][e] + cr[a][b][e], c1[a][b][e] + ci[a][b][e]);
                            out[((int64_t)(m >> 3) * p.n_cols + col) * 8 + (m & 7)] = v;
                            if (p.hdiag && m == col) p.hdiag[(int64_t)cc.rep * p.M + m] = v;
                        }
                    }
            clear_acc();
            cc.t += gridDim.x; load_item(cc);
            if (cc.t < total) warp_range();
        }
    }
}

// ============================================================================ persistent ring GEMM over a column list
// The K2 design (one shared-memory ring of k-tiles running across work items, mbarrier hand-over, three-multiplication
// operand planes, warp-level k clipping) for the two GEMMs of the factored perfect-CSI pass, whose columns are the
// (realization, scheme, SNR) vectors of the whole batch:
//   BGATHER = true   s = G v          A = planes of G (rows = samples), B gathered from the units' interleaved v scratch
//                                      (16-byte cp.async per element, stride b_kstride; the operand sums are formed in registers)
//   BGATHER = false  U = Q^H r        A = planes of Q^H, B = planes of r written by k_apply_h_cols
// EPI = 0: out[col][m] = C;  EPI = 1: the cancellation epilogue y_ic = y - C + h v written into the unit scratch.
// Work items: (row tile, column tile), row tile fastest, so CTAs that run side by side share their B columns in L2.
struct GemmRingParams {
    int M, n_cols, lda, ldb, ldc, n_mt, n_nt;
    const cplx* At1; const double* At2;          // A planes [m][lda] (k-contiguous)
    const cplx* b1; const double* b2;            // B planes [col][ldb]                 (BGATHER = false)
    const cplx* bsrc; const int64_t* b_off; int b_kstride;                               // (BGATHER = true)
    const int* mt_klo; const int* mt_khi; const int* m8_klo; const int* m8_khi;
    cplx* out;
    cplx* e_out; const int64_t* e_off; const cplx* e_y; const int64_t* e_yoff; const cplx* e_h; const int* e_rep;
    const cplx* e_v;
};
template <int WM, int WN, int TMW, bool BGATHER, int EPI>
__global__ void __launch_bounds__(32 * WM * WN, (WM * WN == 8) ? 1 : 2) k_gemm_ring(GemmRingParams p) {
    constexpr int TM = 8 * TMW * WM, TN = 16 * WN, KT = GEMMD_KT, LDS = KT + 4, NTHR = 32 * WM * WN, NS = GEMMD_STAGES;
    extern __shared__ __align__(128) double smem[];
    cplx (*As)[TM][LDS] = reinterpret_cast<cplx (*)[TM][LDS]>(smem);
    cplx (*Bs)[TN][LDS] = reinterpret_cast<cplx (*)[TN][LDS]>(smem + NS * 2 * TM * LDS);
    double (*As2)[TM][LDS] = reinterpret_cast<double (*)[TM][LDS]>(smem + NS * 2 * (TM + TN) * LDS);
    double (*Bs2)[TN][LDS] = reinterpret_cast<double (*)[TN][LDS]>(smem + NS * 2 * (TM + TN) * LDS + NS * TM * LDS);
    __shared__ uint64_t full[NS], empty[NS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t4 = lane & 3;
    if (tid == 0) {
        for (int s_ = 0; s_ < NS; ++s_) { mbar_init(&full[s_], NTHR); mbar_init(&empty[s_], WM * WN); }
    }
    __syncthreads();
    const int total = p.n_mt * p.n_nt;

    struct Cursor { int t, kt, nk, klo, khi, m0, n0; };
    auto load_item = [&](Cursor& c) {
        if (c.t >= total) { c.nk = 0; return; }
        const int nt = c.t / p.n_mt, mt = c.t - nt * p.n_mt;
        c.klo = p.mt_klo[mt] & ~1; c.khi = p.mt_khi[mt];
        c.nk = c.khi > c.klo ? (c.khi - c.klo + KT - 1) / KT : 0;
        c.m0 = mt * TM; c.n0 = nt * TN; c.kt = 0;
    };
    Cursor cc;
    cc.t = blockIdx.x;
    load_item(cc);

    constexpr int EA = (TM * KT) / NTHR, EB = (TN * KT) / NTHR, EA2 = (TM * KT / 2) / NTHR, EB2 = (TN * KT / 2) / NTHR;
    static_assert(EA * NTHR == TM * KT && EB * NTHR == TN * KT && EA2 * NTHR == TM * KT / 2 && EB2 * NTHR == TN * KT / 2, "tile / thread mismatch");
    constexpr int RSTEP = NTHR / KT, RSTEP2 = NTHR / (KT / 2);
    const int kk1 = tid & (KT - 1), r1 = tid / KT, kk2 = (tid & (KT / 2 - 1)) * 2, r2 = tid / (KT / 2);
    const unsigned sA1 = (unsigned)__cvta_generic_to_shared(&As[0][r1][kk1]), sB1 = (unsigned)__cvta_generic_to_shared(&Bs[0][r1][kk1]);
    const unsigned sA2 = (unsigned)__cvta_generic_to_shared(&As2[0][r2][kk2]), sB2 = (unsigned)__cvta_generic_to_shared(&Bs2[0][r2][kk2]);
    constexpr unsigned STG_A1 = TM * LDS * 16, STG_B1 = TN * LDS * 16, STG_A2 = TM * LDS * 8, STG_B2 = TN * LDS * 8;
    int pt = blockIdx.x, pkt = 0, pnk = 0, prem = 0;
    const cplx* ga1 = nullptr; const double* ga2 = nullptr; const double* gb2 = nullptr;
    const cplx* gb1[EB];
    unsigned okA1 = 0, okB1 = 0, okA2 = 0, okB2 = 0;
    auto producer_item = [&]() {
        int nt = 0, mt = 0, klo = 0, khi = 0;
        for (;;) {                                             // items with an empty k-range have no jobs
            if (pt >= total) { pnk = 0; return; }
            nt = pt / p.n_mt; mt = pt - nt * p.n_mt;
            klo = p.mt_klo[mt] & ~1; khi = p.mt_khi[mt];
            if (khi > klo) break;
            pt += gridDim.x;
        }
        pnk = (khi - klo + KT - 1) / KT; pkt = 0; prem = khi - klo;
        const int m0 = mt * TM, n0 = nt * TN;
        ga1 = p.At1 + (int64_t)(m0 + r1) * p.lda + klo + kk1;
        ga2 = p.At2 + (int64_t)(m0 + r2) * p.lda + klo + kk2;
        okA1 = okB1 = okA2 = okB2 = 0;
#pragma unroll
        for (int e = 0; e < EA; ++e) okA1 |= (unsigned)(m0 + r1 + e * RSTEP < p.M) << e;
#pragma unroll
        for (int e = 0; e < EA2; ++e) okA2 |= (unsigned)(m0 + r2 + e * RSTEP2 < p.M) << e;
#pragma unroll
        for (int e = 0; e < EB; ++e) {
            const int col = n0 + r1 + e * RSTEP;
            const bool ok = col < p.n_cols;
            okB1 |= (unsigned)ok << e;
            if (BGATHER) gb1[e] = p.bsrc + (ok ? p.b_off[col] + (int64_t)(klo + kk1) * p.b_kstride : 0);
            else gb1[e] = p.b1 + (ok ? (int64_t)col * p.ldb + klo + kk1 : 0);
        }
        if (!BGATHER) {
            gb2 = p.b2 + (int64_t)(n0 + r2) * p.ldb + klo + kk2;
#pragma unroll
            for (int e = 0; e < EB2; ++e) okB2 |= (unsigned)(n0 + r2 + e * RSTEP2 < p.n_cols) << e;
        }
    };
    auto cp16 = [](unsigned dst, const void* src, bool ok) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(src), "r"(ok ? 16 : 0));
    };
    unsigned pj = 0, cj = 0;
    auto produce = [&]() {
        if (pnk == 0) return;
        const unsigned st = pj % NS;
        if (pj >= NS) mbar_wait(&empty[st], ((pj / NS) - 1) & 1);
        const bool k1 = kk1 < prem, k2 = kk2 < prem;
        const int64_t lda = p.lda, ldb = p.ldb;
#pragma unroll
        for (int e = 0; e < EA; ++e) { const bool ok = k1 && ((okA1 >> e) & 1); cp16(sA1 + st * STG_A1 + e * RSTEP * LDS * 16, ok ? ga1 + e * RSTEP * lda : p.At1, ok); }
#pragma unroll
        for (int e = 0; e < EA2; ++e) { const bool ok = k2 && ((okA2 >> e) & 1); cp16(sA2 + st * STG_A2 + e * RSTEP2 * LDS * 8, ok ? ga2 + e * RSTEP2 * lda : p.At2, ok); }
#pragma unroll
        for (int e = 0; e < EB; ++e) {
            const bool ok = k1 && ((okB1 >> e) & 1);
            cp16(sB1 + st * STG_B1 + e * RSTEP * LDS * 16, ok ? (const void*)gb1[e] : (const void*)p.At1, ok);
            gb1[e] += BGATHER ? (int64_t)KT * p.b_kstride : KT;
        }
        if (!BGATHER) {
#pragma unroll
            for (int e = 0; e < EB2; ++e) { const bool ok = k2 && ((okB2 >> e) & 1); cp16(sB2 + st * STG_B2 + e * RSTEP2 * LDS * 8, ok ? gb2 + e * RSTEP2 * ldb : p.At2, ok); }
            gb2 += KT;
        }
        mbar_arrive_cp_async(&full[st]);
        ++pj;
        ga1 += KT; ga2 += KT; prem -= KT;
        if (++pkt == pnk) { pt += gridDim.x; producer_item(); }
    };
    producer_item();

    double cr[TMW][2][2], ci[TMW][2][2], c1[TMW][2][2];
    auto clear_acc = [&]() {
#pragma unroll
        for (int a = 0; a < TMW; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) { cr[a][b][0] = cr[a][b][1] = ci[a][b][0] = ci[a][b][1] = c1[a][b][0] = c1[a][b][1] = 0.0; }
    };
    int wlo = 0, whi = 0;
    auto warp_range = [&]() {
        wlo = cc.klo; whi = cc.khi;
        if (p.m8_klo) {
            int lo = 0x7fffffff, hi = 0;
#pragma unroll
            for (int x = 0; x < TMW; ++x) {
                const int r = cc.m0 + wm * 8 * TMW + x * 8;
                if (r < p.M) { lo = min(lo, p.m8_klo[r >> 3]); hi = max(hi, p.m8_khi[r >> 3]); }
            }
            wlo = max(wlo, lo); whi = min(whi, hi);
        }
    };
    clear_acc();
    if (cc.t < total) warp_range();
#pragma unroll 1
    for (int i = 0; i < GEMMD_DEPTH; ++i) produce();
#pragma unroll 1
    while (cc.t < total) {
        if (cc.nk > 0) {                                       // (an item with an empty k-range -- e.g. rows inside the zero guard of
            produce();                                         //  an OFDM frame -- has no jobs: the producer must not run further ahead)
            const int st = cj % NS;
            mbar_wait(&full[st], (cj / NS) & 1);
            const int kbase = cc.klo + cc.kt * KT;
#pragma unroll
            for (int kk = 0; kk < KT; kk += 4) {
                if (kbase + kk + 4 <= wlo || kbase + kk >= whi) continue;     // warp-uniform
                cplx a[TMW], b[2];
                double a2[TMW], b2[2];
#pragma unroll
                for (int x = 0; x < TMW; ++x) { a[x] = As[st][wm * 8 * TMW + x * 8 + g][kk + t4]; a2[x] = As2[st][wm * 8 * TMW + x * 8 + g][kk + t4]; }
#pragma unroll
                for (int y = 0; y < 2; ++y) {
                    b[y] = Bs[st][wn * 16 + y * 8 + g][kk + t4];
                    if (BGATHER) { b2[y] = b[y].y - b[y].x; b[y].y = b[y].x + b[y].y; }      // (br, br+bi), bi-br
                    else b2[y] = Bs2[st][wn * 16 + y * 8 + g][kk + t4];
                }
#pragma unroll
                for (int x = 0; x < TMW; ++x)
#pragma unroll
                    for (int y = 0; y < 2; ++y) {
                        dmma884(c1[x][y][0], c1[x][y][1], a[x].y, b[y].x);
                        dmma884(cr[x][y][0], cr[x][y][1], a2[x], b[y].y);
                        dmma884(ci[x][y][0], ci[x][y][1], a[x].x, b2[y]);
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
            ++cj;
            ++cc.kt;
        }
        if (cc.kt >= cc.nk) {
            if (EPI == 1) {
                // y_ic = y - U + h v  (DS.m:541-543 with U = Q^H H G v).  The operands of a row tile's four columns are
                // fetched first (read-only path, so the loads are not ordered behind the stores), then combined and stored.
                int64_t off[2][2], yo[2][2]; int rp[2][2]; bool okc[2][2];
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = cc.n0 + wn * 16 + b * 8 + 2 * t4 + e;
                        okc[b][e] = col < p.n_cols;
                        off[b][e] = okc[b][e] ? p.e_off[col] : 0; yo[b][e] = okc[b][e] ? p.e_yoff[col] : 0; rp[b][e] = okc[b][e] ? p.e_rep[col] : 0;
                    }
#pragma unroll
                for (int a = 0; a < TMW; ++a) {
                    const int m = cc.m0 + wm * 8 * TMW + a * 8 + g;
                    if (m >= p.M) continue;
                    cplx yv[2][2], hv[2][2], vv[2][2];
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            yv[b][e] = ld_nc(p.e_y + yo[b][e] + m);
                            hv[b][e] = ld_nc(p.e_h + (int64_t)rp[b][e] * p.M + m);
                            vv[b][e] = ld_nc(p.e_v + off[b][e] + (int64_t)m * 16);
                        }
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (!okc[b][e]) continue;
                            const cplx hvv = cmul(hv[b][e], vv[b][e]);
                            p.e_out[off[b][e] + (int64_t)m * 16] = cmake(yv[b][e].x - (c1[a][b][e] + cr[a][b][e]) + hvv.x,
                                                                          yv[b][e].y - (c1[a][b][e] + ci[a][b][e]) + hvv.y);
                        }
                }
            } else {
#pragma unroll
                for (int a = 0; a < TMW; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int m = cc.m0 + wm * 8 * TMW + a * 8 + g, col = cc.n0 + wn * 16 + b * 8 + 2 * t4 + e;
                            if (m < p.M && col < p.n_cols)
                                p.out[(int64_t)col * p.ldc + m] = cmake(c1[a][b][e] + cr[a][b][e], c1[a][b][e] + ci[a][b][e]);
                        }
            }
            clear_acc();
            cc.t += gridDim.x; load_item(cc);
            if (cc.t < total) warp_range();
        }
    }
}

// ============================================================================ explicit D-hat (API)
// Dhat[i, i+delta] = sum_p W[i, i+delta, p] hP[p]  scattered into a zeroed dense K x K; diag from W.diag.
__global__ void k_estimate(cplx* __restrict__ Dhat, cplx* __restrict__ hdiag, WTiles w,
                           const int* __restrict__ tile_ptr, const int* __restrict__ tile_delta,
                           const cplx* __restrict__ hP, int K, int P, int P4, int RT) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int n_tiles = tile_ptr[RT];
    if (idx < n_tiles * 8) {
        int t = idx >> 3, r = idx & 7;
        int lo = 0, hi = RT;                      // row tile of t: largest rt with tile_ptr[rt] <= t
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (tile_ptr[mid] <= t) lo = mid; else hi = mid; }
        int i = lo * 8 + r, j = i + tile_delta[t];
        if (i < K && j >= 0 && j < K) {
            cplx acc = cmake(0.0, 0.0);
            for (int p = 0; p < P; ++p) cfma(acc, w.frag[((int64_t)t * P4 + (p >> 2)) * 32 + r * 4 + (p & 3)], hP[p]);
            if (Dhat) Dhat[(int64_t)j * K + i] = acc;
        }
    }
    if (idx < K) {
        cplx acc = cmake(0.0, 0.0);
        for (int p = 0; p < P; ++p) cfma(acc, w.diag[(int64_t)idx * P + p], hP[p]);
        if (Dhat) Dhat[(int64_t)idx * K + idx] = acc;
        if (hdiag) hdiag[idx] = acc;
    }
}

// ============================================================================ K4: fused IC iteration
// One launch per iteration `it` (it = 0 is the one-tap stage).  Every CTA owns up to NC_MAX
// columns and ALL K rows of them:
//   EST  CTA: one scheme, one SNR point, 16 consecutive realizations (they share W_snr)
//   PERF CTA: one realization, all (scheme-on-waveform, SNR) columns (they share D_rep)
// phases: A quantise + precode -> v ; B interference = Woff(hP_prev) v  or  (D - diag h) v (DMMA) ;
//         C new pilot estimates ; D one-tap channel + equalise ; E de-spread / select, demap, count.
// (k_ic_main runs phase B, k_ic_light phases C, D, E and the next iteration's phase A.)
struct IcCta { int mode, scheme_or_wf, snr, first, n_cols; };   // mode 0 EST, 1 PERF
struct IcParams {
    int it, n_iter, n_rep, n_snr, K_max, pilot_rows;      // pilot_rows: rows of the shared pilot tables (4 * max P4)
    int ring_cplx;                                        // complex elements of one warp's W-fragment ring
    int stage_cplx;                                       // complex elements of the main kernel's staging area
    const IcCta* ctas;
    SchemeDev sch[3];
    ConstDev cst[2];
    int wf_scheme[2][2]; int wf_nscheme[2];
    const cplx* D[2];          // [rep][K/8][K][8] true transmission matrices (row-tile-major)
    const int* d_jlo[2]; const int* d_jhi[2];   // per row tile: structural column range of D
    const cplx* htrue[2];      // [rep][K]
    cplx* scratch;             // per CTA: 3 buffers of K_max*NC_MAX
    uint32_t* err;             // [rep][snr][it][scheme][csi][edge]
    unsigned int* queue;       // [n_iter+1] unit counters of the main stage (zeroed per batch)
    int n_units;
    int n_units_main;          // units k_ic_main processes (all of them, or only the EST units in factored mode)
    unsigned long long* trace; // development: per CTA {smid, t0, t_pre, t_main_own, t_main_all, t_end, B busy ns, units}
    double* mse;               // optional: sum_i |h_est[i] - h[i]|^2 per [rep][snr][it][scheme] (estimated-CSI columns); nullptr = off
    // perfect-CSI columns of a waveform whose equalisation / detection / counting is done by k_perfect_fbmc_det: the decided words
    // [column][perf_zw_stride[wf]] (column = rep * nsch * n_snr + slot * n_snr + snr); k_ic_light then only precodes (phase A)
    const uint8_t* perf_zw[2]; int perf_zw_stride[2];
    // schemes whose estimated-CSI cancellation runs in factored form (k_est_factored): k_ic_light leaves h-hat = W_diag hP of the
    // unit's columns in the first scratch buffer, [row][column] like v
    int est_fact_mask;
    // scratch layout of the units the one-column-per-CTA chain kernels read and write (factored EST units, detected PERF units):
    // column-major, [column][K_max] (a column's K values contiguous: coalesced for the chain kernels, 128-byte runs for the DMMA
    // fragments of k_ic_light) -- bit 0: factored EST units, bit 1: detected PERF units; otherwise row-major [row][16] (what
    // k_ic_main / k_ic_est_tc / the GEMM chain use)
    int chain_colmajor;
};

// Column -> (scheme, SNR point, realization); false for an unused slot.  PERF units keep each 8-column half
// (one DMMA n-tile) on one scheme: column c = 8 * (scheme slot on the waveform) + (SNR point - first).
__device__ __forceinline__ bool ic_col(const IcParams& p, const IcCta& c, int col, int& scheme, int& snr, int& rep) {
    if (c.mode == 0) { scheme = c.scheme_or_wf; snr = c.snr; rep = c.first + col; return rep < p.n_rep; }
    const int slot = col >> 3;
    snr = c.first + (col & 7); rep = c.snr;
    if (slot >= p.wf_nscheme[c.scheme_or_wf] || snr >= p.n_snr) { scheme = p.wf_scheme[c.scheme_or_wf][0]; snr = 0; return false; }
    scheme = p.wf_scheme[c.scheme_or_wf][slot];
    return true;
}

#define IC_PILOT_MAX 128
#ifndef EST_WSRC
#define EST_WSRC 0
#endif
#ifndef EST_RING
#define EST_RING 2
#endif        // pilots per scheme supported by the shared pilot tables

__device__ __forceinline__ void cp_async16_plain(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }
// streaming 16-byte load: read-only path, no L1 allocation (W and D fragments are used once per CTA; the
// L1 is kept for the v / y_ic rows, which are re-read many times)
__device__ __forceinline__ cplx ld_stream(const cplx* ptr) {
    cplx r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(ptr));
    return r;
}

// two adjacent complex values (32 bytes, 32-byte aligned) with one 256-bit load
__device__ __forceinline__ void ld_cplx2(const cplx* ptr, cplx& a, cplx& b) {
    asm("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(ptr));
}

__device__ __forceinline__ void st_cplx2(cplx* ptr, cplx a, cplx b) {   // 32 bytes, 32-byte aligned
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(ptr), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

// ---- phase B, estimated CSI: acc[i, c] = sum_{delta} ( sum_p W[i, i+delta, p] hP[p, c] ) * v[i+delta, c]
// One warp owns 8 rows (a row tile) at a time and walks its diagonal tiles.  Per tile the P4 pilot
// quads are DMMA k-steps with the hP fragments as B operand (shared by every tile), followed by an
// element-wise product with v on the C fragment.  The W fragment stream of a row tile is contiguous
// (512 bytes per warp load).
// Tuning knobs (compile-time, see profiles/ for the sweep):
//   EST_WSRC  0: W fragments through a per-lane cp.async ring in shared memory (EST_RING tiles in flight)
//             1: next tile's fragments prefetched into registers with streaming loads
//             2: this tile's fragments loaded where they are used
//   measured (profiles/r01_kic_variant_sweep.txt): 0 is best in the mixed EST+PERF main stage
//   EST_VPRE  1: next tile's v rows prefetched into registers
//   EST_HPREG 1: the hP B-fragments live in registers for the whole row tile (P4 <= 4 only); always on in the
//             EST-only kernel instance of the factored mode (template flag HPR), where they fit without spills
#define EST_H1S 20           // row stride (doubles) of the real-part plane of the pilot estimates
#ifndef EST_VPRE
#define EST_VPRE 0
#endif
#ifndef EST_HPREG
#define EST_HPREG 0
#endif
#ifndef EST_RING
#define EST_RING 2
#endif
template <int P4T, bool VREAL, bool HPR>
__device__ IC_INLINE void est_interference(const cplx* __restrict__ frag, const int* __restrict__ tptr,
                                                 const int* __restrict__ tdel, const cplx* hPs, const cplx* hP3, const double* hP1,
                                                 const cplx* vbuf,
                                                 cplx* ybuf, const cplx* const* ycolp, cplx* ring, int K, int warp,
                                                 int nwarp, int lane) {
    constexpr int NC = NC_MAX, HS = NC + 2, ST = EST_RING;
    constexpr bool HPREG = (EST_HPREG || HPR) && P4T <= 4;
    // hP1: real parts of the pilot estimates as a plane of doubles, row stride 20 -> the 8-byte fragment loads
    // are conflict-free (the 16-byte stride of the complex table would give 2-way conflicts)
    const int g = lane >> 2, t4 = lane & 3;
    const int RT = (K + 7) / 8;
    cplx* slot = ring + lane;                                  // [stage][pq][32 lanes]
    cplx hb[HPREG ? P4T : 1][2], hb3[HPREG && CHEST_3M ? P4T : 1][2];
    if (HPREG) {
#pragma unroll
        for (int pq = 0; pq < P4T; ++pq)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                hb[HPREG ? pq : 0][ct] = hPs[(pq * 4 + t4) * HS + ct * 8 + g];
                if (CHEST_3M) hb3[HPREG && CHEST_3M ? pq : 0][ct] = hP3[(pq * 4 + t4) * HS + ct * 8 + g];
            }
    }
    for (int rt = warp; rt < RT; rt += nwarp) {
        double accr[2][2] = {{0, 0}, {0, 0}}, acci[2][2] = {{0, 0}, {0, 0}};
        const int i = rt * 8 + g;
        const int t0 = tptr[rt], tend = tptr[rt + 1];
        cplx nxt[EST_WSRC == 1 ? P4T : 1], vn[2][2];
        auto load_w = [&](int t, int s) {
            if (EST_WSRC == 0) {
                if (t < tend) {
#pragma unroll
                    for (int pq = 0; pq < P4T; ++pq)
                        cp_async16_plain(slot + (s * P4T + pq) * 32, frag + ((int64_t)t * P4T + pq) * 32 + lane);
                }
                cp_async_commit();
            } else if (t < tend) {
#pragma unroll
                for (int pq = 0; pq < P4T; ++pq) nxt[EST_WSRC == 1 ? pq : 0] = ld_stream(frag + ((int64_t)t * P4T + pq) * 32 + lane);
            }
        };
        auto load_v = [&](int t, cplx (&dst)[2][2]) {
            int j = i + tdel[t];
            j = j < 0 ? 0 : (j > K - 1 ? K - 1 : j);
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) ld_cplx2(vbuf + j * NC + ct * 8 + 2 * t4, dst[ct][0], dst[ct][1]);
        };
        if (EST_WSRC == 0) {
#pragma unroll
            for (int s = 0; s < ST; ++s) load_w(t0 + s, s);
        } else if (EST_WSRC == 1) {
            load_w(t0, 0);
        }
        if (EST_VPRE && t0 < tend) load_v(t0, vn);
        for (int t = t0; t < tend; ++t) {
            const int s = (t - t0) % ST;
            cplx cur[P4T], v[2][2];
            if (EST_VPRE) {
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) v[ct][e] = vn[ct][e];
                if (t + 1 < tend) load_v(t + 1, vn);
            } else {
                load_v(t, v);
            }
            if (EST_WSRC == 0) {
                cp_async_wait_group<ST - 1>();
#pragma unroll
                for (int pq = 0; pq < P4T; ++pq) cur[pq] = slot[(s * P4T + pq) * 32];
            } else if (EST_WSRC == 1) {
#pragma unroll
                for (int pq = 0; pq < P4T; ++pq) cur[pq] = nxt[EST_WSRC == 1 ? pq : 0];
                load_w(t + 1, 0);
            } else {                                          // direct streaming loads of this tile's fragments
#pragma unroll
                for (int pq = 0; pq < P4T; ++pq) cur[pq] = ld_stream(frag + ((int64_t)t * P4T + pq) * 32 + lane);
            }
            double tr[2][2] = {{0, 0}, {0, 0}}, ti[2][2] = {{0, 0}, {0, 0}};
#if CHEST_3M
            // three-multiplication complex product: hP3 holds (-(br+bi), bi-br) of the pilot estimates.  The common
            // term t1 = sum_p (ar+ai) br is accumulated first; the re and im chains then START from it (accumulator
            // input of their first DMMA), so no separate additions are needed
            double t1[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
            for (int pq = 0; pq < P4T; ++pq) {
                const double as = cur[pq].x + cur[pq].y;
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) {
                    const double b1 = HPREG ? hb[HPREG ? pq : 0][ct].x : hP1[(pq * 4 + t4) * EST_H1S + ct * 8 + g];
                    dmma884(t1[ct][0], t1[ct][1], as, b1);
                }
            }
#pragma unroll
            for (int pq = 0; pq < P4T; ++pq) {
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) {
                    const cplx b23 = HPREG ? hb3[HPREG ? pq : 0][ct] : hP3[(pq * 4 + t4) * HS + ct * 8 + g];
                    if (pq == 0) {
                        dmma884c(tr[ct][0], tr[ct][1], cur[pq].y, b23.x, t1[ct][0], t1[ct][1]);
                        dmma884c(ti[ct][0], ti[ct][1], cur[pq].x, b23.y, t1[ct][0], t1[ct][1]);
                    } else {
                        dmma884(tr[ct][0], tr[ct][1], cur[pq].y, b23.x);
                        dmma884(ti[ct][0], ti[ct][1], cur[pq].x, b23.y);
                    }
                }
            }
#else
#pragma unroll
            for (int pq = 0; pq < P4T; ++pq) {
                const double nai = dneg(cur[pq].y);
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) {
                    const cplx b = HPREG ? hb[HPREG ? pq : 0][ct] : hPs[(pq * 4 + t4) * HS + ct * 8 + g];
                    dmma884(tr[ct][0], tr[ct][1], cur[pq].x, b.x);
                    dmma884(tr[ct][0], tr[ct][1], nai, b.y);
                    dmma884(ti[ct][0], ti[ct][1], cur[pq].x, b.y);
                    dmma884(ti[ct][0], ti[ct][1], cur[pq].y, b.x);
                }
            }
#endif
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (VREAL) {                               // v exactly real: T o v needs half the multiply-adds
                        accr[ct][e] = fma(tr[ct][e], v[ct][e].x, accr[ct][e]);
                        acci[ct][e] = fma(ti[ct][e], v[ct][e].x, acci[ct][e]);
                    } else {
                        accr[ct][e] = fma(tr[ct][e], v[ct][e].x, accr[ct][e]); accr[ct][e] = fma(-ti[ct][e], v[ct][e].y, accr[ct][e]);
                        acci[ct][e] = fma(tr[ct][e], v[ct][e].y, acci[ct][e]); acci[ct][e] = fma(ti[ct][e], v[ct][e].x, acci[ct][e]);
                    }
                }
            if (EST_WSRC == 0) load_w(t + ST, s);              // refill the slot just consumed
        }
        if (EST_WSRC == 0) cp_async_wait_group<0>();
        if (i < K) {
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int c = ct * 8 + 2 * t4 + e;
                    const cplx* yp = ycolp[c];
                    cplx yv = yp ? yp[i] : cmake(0.0, 0.0);
                    ybuf[i * NC + c] = cmake(yv.x - accr[ct][e], yv.y - acci[ct][e]);
                }
        }
    }
}

// generic pilot count (P4 not 4 or 8): direct fragment loads, no staging
__device__ IC_INLINE void est_interference_generic(const cplx* __restrict__ frag, const int* __restrict__ tptr,
                                                         const int* __restrict__ tdel, const cplx* hPs,
                                                         const cplx* vbuf, cplx* ybuf, const cplx* const* ycolp, int K,
                                                         int P4, int warp, int nwarp, int lane) {
    constexpr int NC = NC_MAX, HS = NC + 2;
    const int g = lane >> 2, t4 = lane & 3;
    const int RT = (K + 7) / 8;
    for (int rt = warp; rt < RT; rt += nwarp) {
        double accr[2][2] = {{0, 0}, {0, 0}}, acci[2][2] = {{0, 0}, {0, 0}};
        const int i = rt * 8 + g;
        for (int t = tptr[rt]; t < tptr[rt + 1]; ++t) {
            double tr[2][2] = {{0, 0}, {0, 0}}, ti[2][2] = {{0, 0}, {0, 0}};
            const cplx* fr = frag + (int64_t)t * P4 * 32 + lane;
            for (int pq = 0; pq < P4; ++pq) {
                cplx a = __ldg(fr + pq * 32);
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) {
                    cplx b = hPs[(pq * 4 + t4) * HS + ct * 8 + g];
                    zmma884(tr[ct], ti[ct], a.x, a.y, b.x, b.y, dneg(b.y));
                }
            }
            int j = i + tdel[t];
            j = j < 0 ? 0 : (j > K - 1 ? K - 1 : j);
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    cplx v = vbuf[j * NC + ct * 8 + 2 * t4 + e];
                    accr[ct][e] = fma(tr[ct][e], v.x, accr[ct][e]); accr[ct][e] = fma(-ti[ct][e], v.y, accr[ct][e]);
                    acci[ct][e] = fma(tr[ct][e], v.y, acci[ct][e]); acci[ct][e] = fma(ti[ct][e], v.x, acci[ct][e]);
                }
        }
        if (i < K) {
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int c = ct * 8 + 2 * t4 + e;
                    const cplx* yp = ycolp[c];
                    cplx yv = yp ? yp[i] : cmake(0.0, 0.0);
                    ybuf[i * NC + c] = cmake(yv.x - accr[ct][e], yv.y - acci[ct][e]);
                }
        }
    }
}

// ---- phase B, perfect CSI: acc[i, c] = sum_{j != i} D[i, j] v[j, c]   (DS.m:541-543)
// A mini-GEMM per CTA: the warps take one row tile each (a "row block" of nwarp row tiles) and walk the
// columns j together in chunks of 32.  The v chunk (B operand) is staged ONCE per chunk in shared memory
// for all warps, two stages, through registers: the loads of chunk ch+1 are in flight while chunk ch is
// multiplied, and the store converts v into the three-multiplication operands  b1 = vr (plane v1, row
// stride 20 doubles) and (b2, b3) = (-(vr+vi), vi-vr) (plane v23, row stride 18 complex) -- both strides
// make the fragment loads conflict-free.  Each warp streams its own row tile of the row-tile-major D
// (A operand: 512 contiguous bytes per k-step) through registers, 16 columns ahead.
#define PERF_CHUNK 32          // v rows staged per barrier
#define PERF_DSUB 16           // D columns prefetched in registers at a time
#define PERF_V1S 20            // row stride (doubles) of the b1 plane
#ifndef IC_THREADS
#define IC_THREADS 256
#endif
#define PERF_STAGE_CPLX (PERF_CHUNK * (NC_MAX + 2) + PERF_CHUNK * PERF_V1S / 2)   // one stage, in complex elements
// REALMASK bit ct: the v values of n-tile ct are exactly real (real precoder and constellation, e.g. the
// data-spreading scheme): the complex x real product needs two DMMAs (ar v, ai v) instead of three.
template <int NCT, int REALMASK>
__device__ IC_INLINE void perf_interference(const cplx* __restrict__ Dm, const cplx* vbuf, cplx* ybuf,
                                                  const cplx* const* ycolp, cplx* vs, const int* __restrict__ rt_jlo,
                                                  const int* __restrict__ rt_jhi, int K, int warp, int nwarp,
                                                  int lane, int tid) {
    constexpr int NC = NC_MAX, U = PERF_DSUB / 4, NSUB = PERF_CHUNK / PERF_DSUB, VS = NC + 2;
    constexpr int EPT = (PERF_CHUNK * NC + IC_THREADS - 1) / IC_THREADS;
    const int g = lane >> 2, t4 = lane & 3;
    const int RT = (K + 7) / 8;
    const int nblk = (RT + nwarp - 1) / nwarp;
    cplx* v23 = vs;                                                            // [stage][row][VS]
    double* v1 = reinterpret_cast<double*>(vs + 2 * PERF_CHUNK * VS);          // [stage][row][PERF_V1S]
    for (int blk = 0; blk < nblk; ++blk) {
        // columns outside [jlo, jhi) are structurally zero for every row of this block (no support overlap
        // between Q_i and H G_j): they are skipped exactly
        int jlo = K, jhi = 0;
        for (int r = blk * nwarp; r < min(RT, (blk + 1) * nwarp); ++r) { jlo = min(jlo, rt_jlo[r]); jhi = max(jhi, rt_jhi[r]); }
        const int ch0 = jlo / PERF_CHUNK, nchunk = jhi > jlo ? (jhi + PERF_CHUNK - 1) / PERF_CHUNK : ch0;
        const int rt = blk * nwarp + warp;
        const bool active = rt < RT;
        double accr[2][2] = {{0, 0}, {0, 0}}, acci[2][2] = {{0, 0}, {0, 0}}, acc1[2][2] = {{0, 0}, {0, 0}};
        const int i = rt * 8 + g;
        const cplx* Drt = Dm + (int64_t)(active ? rt : 0) * K * 8 + g;
        cplx an[U], vreg[EPT];
        auto load_d = [&](int j0) {                            // 16 columns of this warp's row tile -> registers
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int j = j0 + 4 * u + t4;
                j = j < K ? j : K - 1;                         // out-of-range columns are masked at use
                an[u] = active ? ld_stream(Drt + (int64_t)j * 8) : cmake(0.0, 0.0);
            }
        };
        auto fetch_v = [&](int ch) {                           // v chunk -> registers (shared by all warps)
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int idx = tid + e * IC_THREADS, jr = idx / NC, c = idx % NC;
                int j = ch * PERF_CHUNK + jr;
                j = j < K ? j : K - 1;
                vreg[e] = vbuf[j * NC + c];
            }
        };
        auto store_v = [&](int s) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int idx = tid + e * IC_THREADS, jr = idx / NC, c = idx % NC;
                if (idx < PERF_CHUNK * NC) {
                    const cplx b = vreg[e];
                    v1[(s * PERF_CHUNK + jr) * PERF_V1S + c] = b.x;
                    v23[(s * PERF_CHUNK + jr) * VS + c] = CHEST_3M ? cmake(-(b.x + b.y), b.y - b.x) : cmake(b.y, 0.0);
                }
            }
        };
        __syncthreads();                                       // previous row block done with both stages
        if (ch0 < nchunk) { fetch_v(ch0); load_d(ch0 * PERF_CHUNK); store_v(ch0 & 1); }
        __syncthreads();
        for (int ch = ch0; ch < nchunk; ++ch) {
            const int s = ch & 1;
            if (ch + 1 < nchunk) fetch_v(ch + 1);
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                cplx a[U];
#pragma unroll
                for (int u = 0; u < U; ++u) a[u] = an[u];
                const int jn = ch * PERF_CHUNK + (sub + 1) * PERF_DSUB;
                if (jn < nchunk * PERF_CHUNK) load_d(jn);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    int j = ch * PERF_CHUNK + sub * PERF_DSUB + 4 * u + t4;
                    if (i >= K || j >= K || i == j) a[u] = cmake(0.0, 0.0);
                    const int row = s * PERF_CHUNK + sub * PERF_DSUB + 4 * u + t4;
#if CHEST_3M
                    const double as = a[u].x + a[u].y;       // dead code when every n-tile is real
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        const double b1 = v1[row * PERF_V1S + ct * 8 + g];
                        if ((REALMASK >> ct) & 1) {
                            dmma884(accr[ct][0], accr[ct][1], a[u].x, b1);
                            dmma884(acci[ct][0], acci[ct][1], a[u].y, b1);
                        } else {
                            const cplx b23 = v23[row * VS + ct * 8 + g];
                            dmma884(acc1[ct][0], acc1[ct][1], as, b1);
                            dmma884(accr[ct][0], accr[ct][1], a[u].y, b23.x);
                            dmma884(acci[ct][0], acci[ct][1], a[u].x, b23.y);
                        }
                    }
#else
                    const double nai = dneg(a[u].y);
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        const double bx = v1[row * PERF_V1S + ct * 8 + g], by = v23[row * VS + ct * 8 + g].x;
                        dmma884(accr[ct][0], accr[ct][1], a[u].x, bx);
                        dmma884(accr[ct][0], accr[ct][1], nai, by);
                        dmma884(acci[ct][0], acci[ct][1], a[u].x, by);
                        dmma884(acci[ct][0], acci[ct][1], a[u].y, bx);
                    }
#endif
                }
            }
            if (ch + 1 < nchunk) store_v(s ^ 1);               // stage s^1 was last read before the previous barrier
            __syncthreads();                                   // chunk ch+1 is staged, chunk ch is consumed
        }
        if (active && i < K) {
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int c = ct * 8 + 2 * t4 + e;
                    const cplx* yp = ycolp[c];
                    cplx yv = yp ? yp[i] : cmake(0.0, 0.0);
                    ybuf[i * NC + c] = cmake(yv.x - (accr[ct][e] + acc1[ct][e]), yv.y - (acci[ct][e] + acc1[ct][e]));
                }
        }
    }
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define IC_TRACE(k, v) do { if (p.trace && threadIdx.x == 0) p.trace[(int64_t)blockIdx.x * 8 + (k)] = (v); } while (0)
// ---------------------------------------------------------------------------------------------
// K4: one interference-cancellation iteration = two launches over the same work units (IcCta):
//   EST  unit: one scheme, one SNR point, 16 consecutive realizations (they share W_snr)
//   PERF unit: one realization, all (scheme-on-waveform, SNR) columns (they share D_rep)
// each owning up to 16 columns and ALL K rows of them.
//   k_ic_main(it)   phase B: y_ic = y - Woff(hP_prev) v   or   y - (D - diag h) v   -- FP64 tensor pipe (DMMA);
//                   persistent CTAs (2 per SM) pull units from a queue ordered heavy-first
//   k_ic_light(it)  phases C, D, E of iteration it (pilot estimates, one-tap channel, equalise, de-spread /
//                   select, demap, count) followed by phase A of iteration it+1 (z = [xP; quantised symbols],
//                   v = C z) -- scalar FP64, latency-bound, so it runs at high occupancy and never shares an SM
//                   with the tensor phase (both use the SM's one FP64 pipe: a DFMA queued behind another CTA's
//                   DMMAs takes ~50 cycles).  it = 0 is the one-tap stage (DS.m:406-433).
// ---------------------------------------------------------------------------------------------
struct IcShared {
    unsigned int cnt[NC_MAX][2];
    int c_scheme[NC_MAX], c_snr[NC_MAX], c_rep[NC_MAX];
    const cplx* ycolp[NC_MAX];
    ConstDev cst[2];
    int unit;
    double mse[NC_MAX];
};

__device__ __forceinline__ void ic_load_unit(const IcParams& p, const IcCta& cta, IcShared& sh) {
    const int tid = threadIdx.x;
    __syncthreads();                                   // previous unit's readers are done
    if (tid < NC_MAX) {
        int s_ = 0, n_ = 0, r_ = 0;
        const bool ok = tid < cta.n_cols && ic_col(p, cta, tid, s_, n_, r_);
        sh.c_scheme[tid] = s_; sh.c_snr[tid] = n_; sh.c_rep[tid] = ok ? r_ : -1;
        sh.ycolp[tid] = ok ? p.sch[s_].y + ((int64_t)n_ * p.n_rep + r_) * p.sch[s_].K : nullptr;
        sh.cnt[tid][0] = sh.cnt[tid][1] = 0;
        sh.mse[tid] = 0.0;
    }
    __syncthreads();
}

#ifndef IC_MIN_BLOCKS
#define IC_MIN_BLOCKS 2
#endif
// The loop bodies compiled into one instance are chosen by the host for the configuration at hand (more
// bodies in one kernel cost registers in all of them): P4S = pilot-quad count of the specialised estimated-CSI
// loop (other schemes take the generic loop), M2 / M1 = real-v mask every 16-column / 8-column perfect-CSI unit
// of the configuration satisfies (0 is always valid).
template <int P4S, int M2, int M1>
__global__ void __launch_bounds__(IC_THREADS, IC_MIN_BLOCKS) k_ic_main(IcParams p) {
    constexpr int NC = NC_MAX, HS = NC + 2;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int it = p.it;
    extern __shared__ __align__(128) cplx ic_smem[];
    // staging area (first, 128-byte aligned), one unit at a time: the v chunks of a PERF unit (2 stages) or
    // the per-warp W-fragment rings of an EST unit -- never both; then the pilot tables of an EST unit
    cplx* vstage = ic_smem;
    cplx* hPs = ic_smem + p.stage_cplx;                     // previous pilot estimates [p][col]
    cplx* hP3 = hPs + p.pilot_rows * HS;                    // their three-multiplication operands (-(re+im), im-re)
    double* hP1 = reinterpret_cast<double*>(hP3 + p.pilot_rows * HS);   // real parts, [p][EST_H1S]
    cplx* ring = vstage + warp * p.ring_cplx;
    __shared__ IcShared sh;
    IC_TRACE(0, smid()); IC_TRACE(1, gtime());
    unsigned long long n_done = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) sh.unit = (int)atomicAdd(p.queue + it, 1u);
        __syncthreads();
        const int unit = sh.unit;
        if (unit >= p.n_units_main) break;
        ++n_done;
        const IcCta cta = p.ctas[unit];
        ic_load_unit(p, cta, sh);
        const int csi = cta.mode;
        const int wf = (cta.mode == 0) ? p.sch[cta.scheme_or_wf].waveform : cta.scheme_or_wf;
        const int K = p.sch[p.wf_scheme[wf][0]].K;
        cplx* vbuf = p.scratch + (int64_t)unit * 3 * p.K_max * NC + (int64_t)p.K_max * NC;
        cplx* ybuf = vbuf + (int64_t)p.K_max * NC;
        if (csi == 0) {
            const SchemeDev& sd = p.sch[cta.scheme_or_wf];
            for (int idx = tid; idx < sd.P4 * 4 * NC; idx += nthr) {      // rows P..4*P4-1 are zero padding
                int c = idx % NC, pp = idx / NC;
                const cplx b = (sh.c_rep[c] >= 0 && pp < sd.P)
                                   ? sd.hP[((int64_t)sh.c_snr[c] * p.n_rep + sh.c_rep[c]) * sd.P + pp] : cmake(0.0, 0.0);
                hPs[pp * HS + c] = b;
                hP3[pp * HS + c] = cmake(-(b.x + b.y), b.y - b.x);
                hP1[pp * EST_H1S + c] = b.x;
            }
            __syncthreads();
            // the D-hat being cancelled is the one estimated in iteration it-1 (DS.m:475,492)
            const int var_prev = (it - 1 == 0 || (it - 1) <= p.n_iter / 2) ? 0 : 1;
            // the EST-only instance (factored mode) has the registers to keep the pilot operands resident
#define EST_CALL(P4T_, VR_) est_interference<P4T_, VR_, (M2 < 0)>(sd.w[var_prev][cta.snr].frag, sd.tile_ptr[var_prev], sd.tile_delta[var_prev], \
                                                        hPs, hP3, hP1, vbuf, ybuf, sh.ycolp, ring, K, warp, nwarp, lane)
            if (sd.P4 == P4S) { if (sd.v_real) EST_CALL(P4S, true); else EST_CALL(P4S, false); }
            else
                est_interference_generic(sd.w[var_prev][cta.snr].frag, sd.tile_ptr[var_prev], sd.tile_delta[var_prev],
                                         hPs, vbuf, ybuf, sh.ycolp, K, sd.P4, warp, nwarp, lane);
        } else {
            const cplx* Dm = p.D[wf] + (int64_t)cta.snr * (((K + 7) / 8) * 8) * K;     // cta.snr holds the realization
            // each 8-column half is on one scheme: real v of a half -> two DMMAs per product instead of three
            // (the masks M2 / M1 hold for every unit of this configuration, see the host)
#define PERF_CALL(NCT_, MASK_) perf_interference<NCT_, MASK_>(Dm, vbuf, ybuf, sh.ycolp, vstage, p.d_jlo[wf], p.d_jhi[wf], K, warp, nwarp, lane, tid)
            if (M2 < 0) {}                                     // EST-only instance (factored mode): no perfect-CSI bodies
            else if (cta.n_cols <= 8) PERF_CALL(1, (M1 < 0 ? 0 : M1));
            else PERF_CALL(2, (M2 < 0 ? 0 : M2));
#undef PERF_CALL
        }
    }
    IC_TRACE(3, gtime()); IC_TRACE(7, n_done);
}

#ifndef IC_LIGHT_BLOCKS
#define IC_LIGHT_BLOCKS 4
#endif
#ifndef IC_LIGHT_WAVES
#define IC_LIGHT_WAVES 64      // grid of the light kernel in units of one resident wave (i.e. one CTA per unit)
#endif
#ifndef IC_LIGHT_THREADS
#define IC_LIGHT_THREADS 256
#endif
#ifndef IC_LIGHT_UNROLL
#define IC_LIGHT_UNROLL 4      // rows in flight per thread in the perfect-CSI equalisation pass (8 spills at 64 registers)
#endif
// Hard decision of one data-symbol estimate: bit errors against the transmitted word, and the decided word
// (the quantised symbol of the next iteration's cancellation, DS.m:482-484).
__device__ __forceinline__ int ic_decide(const ConstDev& cd, cplx xd, uint32_t tw, uint32_t em, unsigned& e_all, unsigned& e_edge) {
    const int word = demap_word(cd, xd);
    const uint32_t diff = (uint32_t)word ^ tw;
    e_all += __popc(diff);
    e_edge += __popc(diff & em);
    return word;
}
__global__ void __launch_bounds__(IC_LIGHT_THREADS, IC_LIGHT_BLOCKS) k_ic_light(IcParams p) {
    constexpr int NC = NC_MAX, HS = NC + 2;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int it = p.it;
    extern __shared__ __align__(128) cplx ic_smem[];
    cplx* hPn = ic_smem;                                    // new pilot estimates [p][col] (zero padded to 4*P4 rows)
    cplx* xPs = hPn + p.pilot_rows * HS;                    // transmitted pilots of the unit's columns [p][col]
    __shared__ IcShared sh;
    uint8_t* zw;                                            // decided words of the data symbols [d][col]
    {   // constellation tables (levels, grid -> word, word -> symbol) in shared memory
        cplx* q = xPs + p.pilot_rows * NC;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const ConstDev& cg = p.cst[k];
            cplx* sym = q; q += cg.order;
            double* lev = reinterpret_cast<double*>(q); q += (cg.n_axis + 1) / 2;
            int* gr = reinterpret_cast<int*>(q); q += (cg.order + 3) / 4;
            for (int e = tid; e < cg.order; e += nthr) { sym[e] = cg.symbol[e]; gr[e] = cg.word_of_grid[e]; }
            for (int e = tid; e < cg.n_axis; e += nthr) lev[e] = cg.level[e];
            if (tid == 0) { sh.cst[k] = cg; sh.cst[k].symbol = sym; sh.cst[k].level = lev; sh.cst[k].word_of_grid = gr; }
        }
        zw = reinterpret_cast<uint8_t*>(q);
    }
    const bool last = it == p.n_iter;                       // the last iteration leaves the state chest_get_state reads
    const bool next_pre = !last;                            // otherwise build v of iteration it+1
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const IcCta cta = p.ctas[unit];
        ic_load_unit(p, cta, sh);
        const int csi = cta.mode;
        const int wf = (cta.mode == 0) ? p.sch[cta.scheme_or_wf].waveform : cta.scheme_or_wf;
        const int K = p.sch[p.wf_scheme[wf][0]].K;
        cplx* vbuf = p.scratch + (int64_t)unit * 3 * p.K_max * NC + (int64_t)p.K_max * NC;
        const cplx* ybuf = vbuf + (int64_t)p.K_max * NC;
        const bool detected = csi == 1 && p.perf_zw[wf] != nullptr;      // equalised, decided and counted by k_perfect_fbmc_det
        if (detected && !next_pre) continue;
        // units whose v / y_ic / h-hat go through the chain kernels keep them column-major (IcParams::chain_colmajor)
        const bool cm = ((p.chain_colmajor & 2) && detected) || ((p.chain_colmajor & 1) && csi == 0 && ((p.est_fact_mask >> cta.scheme_or_wf) & 1));
        const int KM = p.K_max;
        auto VI = [&](int i, int c) -> int64_t { return cm ? (int64_t)c * KM + i : (int64_t)i * NC + c; };
        auto ZI = [&](int d, int c) -> int { return cm ? c * KM + d : d * NC + c; };      // the decided words follow the layout
        // y_ic of this unit: the cancelled symbols, or y itself in the one-tap stage
        auto yic = [&](int i, int c) -> cplx {
            if (it > 0) return ybuf[VI(i, c)];
            const cplx* yp = sh.ycolp[c];
            return yp ? yp[i] : cmake(1.0, 0.0);
        };
        bool any_despread = false;
        for (int c = 0; c < cta.n_cols; ++c) any_despread |= p.sch[sh.c_scheme[c]].detect_mode == 1;
        if (detected) {
            any_despread = false;
            const int nv = p.wf_nscheme[wf] * p.n_snr;
            // sixteen decided words of one column per 16-byte load (the column stride is a multiple of 16 bytes)
            for (int idx = tid; idx < ((p.K_max + 15) >> 4) * NC; idx += nthr) {
                const int c = idx % NC, d0 = (idx / NC) << 4;
                if (sh.c_rep[c] < 0) continue;
                const int n_data = p.sch[sh.c_scheme[c]].n_data;
                if (d0 >= n_data) continue;
                const int64_t col = (int64_t)sh.c_rep[c] * nv + (c >> 3) * p.n_snr + sh.c_snr[c];
                const uint4 w = *reinterpret_cast<const uint4*>(p.perf_zw[wf] + col * p.perf_zw_stride[wf] + d0);
                if (cm && (KM & 15) == 0) { *reinterpret_cast<uint4*>(zw + c * KM + d0) = w; continue; }      // same order in shared memory
                const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (d0 + j < n_data) zw[ZI(d0 + j, c)] = (uint8_t)(ww[j >> 2] >> (8 * (j & 3)));
            }
        }
        {   // transmitted pilots of the columns (phase C divides by them, phase A re-inserts them)
            const int Pw = p.sch[p.wf_scheme[wf][0]].P;
            for (int idx = tid; idx < Pw * NC; idx += nthr) {
                const int c = idx % NC, pp = idx / NC;
                xPs[pp * NC + c] = sh.c_rep[c] >= 0 ? p.sch[sh.c_scheme[c]].xP[(int64_t)sh.c_rep[c] * Pw + pp] : cmake(1.0, 0.0);
            }
            __syncthreads();
        }
        // ---- phase C: LS pilot estimates from the (cancelled) symbols   (DS.m:412-414, 487-489)
        if (csi == 0) {
            const SchemeDev& sd = p.sch[cta.scheme_or_wf];
            for (int idx = tid; idx < sd.P4 * 4 * NC; idx += nthr) {      // rows P..4*P4-1: zero padding for the DMMA
                int c = idx % NC, pp = idx / NC;
                cplx hp = cmake(0.0, 0.0);
                if (sh.c_rep[c] >= 0 && pp < sd.P) {
                    cplx q = cdiv(yic(sd.pilot_pos[pp], c), xPs[pp * NC + c]);
                    hp = cmake(q.x / sd.sqrt_kappa, q.y / sd.sqrt_kappa);
                    sd.hP[((int64_t)sh.c_snr[c] * p.n_rep + sh.c_rep[c]) * sd.P + pp] = hp;
                }
                hPn[pp * HS + c] = hp;
            }
            __syncthreads();
        }
        // ---- phases D + E: one-tap channel, equalisation (DS.m:428-429, 515-521) and, for the schemes whose data
        //      symbols are read off single positions, the decision right away (DS.m:430-433): bit errors and the
        //      decided word.  De-spread schemes write the equalised symbols to vbuf and are decided below.
        {
            const int var_cur = (it == 0 || it <= p.n_iter / 2) ? 0 : 1;
            if (csi == 0) {
                // h_est = W_diag * hP as a small DMMA product per row tile (A = W[i,i,p] fragments, B = hP_new),
                // then x_hat = y_ic / h_est on the C fragment
                const SchemeDev& sd = p.sch[cta.scheme_or_wf];
                const ConstDev& cd = sh.cst[sd.constellation];
                const int g = lane >> 2, t4 = lane & 3, RT = (K + 7) / 8, P4 = sd.P4;
                const bool select = sd.detect_mode != 1;
                const cplx* __restrict__ wf_ = sd.wdiag_frag[var_cur] + (int64_t)cta.snr * RT * P4 * 32;
                const bool hfact = next_pre && ((p.est_fact_mask >> cta.scheme_or_wf) & 1);     // k_est_factored adds h-hat v back
                cplx* hbuf = vbuf - (int64_t)p.K_max * NC;
                unsigned e_all[2][2] = {{0, 0}, {0, 0}}, e_edge[2][2] = {{0, 0}, {0, 0}};
                double dm[2][2] = {{0, 0}, {0, 0}};                // channel-estimation error of this lane's columns (p.mse)
                for (int rt = warp; rt < RT; rt += nwarp) {
                    double hr[2][2] = {{0, 0}, {0, 0}}, hi[2][2] = {{0, 0}, {0, 0}};
                    const int i = rt * 8 + g;
                    cplx yv[2][2];
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct) {
                        if (it > 0 && i < K && !cm) ld_cplx2(ybuf + i * NC + ct * 8 + 2 * t4, yv[ct][0], yv[ct][1]);      // the lane's column pair: 32 bytes
                        else
#pragma unroll
                            for (int e = 0; e < 2; ++e) yv[ct][e] = i < K ? yic(i, ct * 8 + 2 * t4 + e) : cmake(0.0, 0.0);
                    }
                    const int d = (select && i < K) ? sd.pos2data[i] : -1;
                    const uint32_t em = d >= 0 ? sd.edge_mask[d] : 0;
                    for (int pq0 = 0; pq0 < P4; pq0 += 4) {         // four diagonal-W fragments in flight (the loads were serialised)
                        cplx a4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            a4[u] = pq0 + u < P4 ? ld_nc(wf_ + ((int64_t)rt * P4 + pq0 + u) * 32 + lane) : cmake(0.0, 0.0);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (pq0 + u >= P4) break;
                            const cplx a = a4[u];
                            const double nai = dneg(a.y);
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct) {
                                const cplx b = hPn[((pq0 + u) * 4 + t4) * HS + ct * 8 + g];
                                dmma884(hr[ct][0], hr[ct][1], a.x, b.x);
                                dmma884(hr[ct][0], hr[ct][1], nai, b.y);
                                dmma884(hi[ct][0], hi[ct][1], a.x, b.y);
                                dmma884(hi[ct][0], hi[ct][1], a.y, b.x);
                            }
                        }
                    }
                    if (i < K) {
                        if (hfact)
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct) {
                                if (cm) {
                                    hbuf[VI(i, ct * 8 + 2 * t4)] = cmake(hr[ct][0], hi[ct][0]);
                                    hbuf[VI(i, ct * 8 + 2 * t4 + 1)] = cmake(hr[ct][1], hi[ct][1]);
                                } else st_cplx2(hbuf + i * NC + ct * 8 + 2 * t4, cmake(hr[ct][0], hi[ct][0]), cmake(hr[ct][1], hi[ct][1]));
                            }
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = ct * 8 + 2 * t4 + e;
                                if (sh.c_rep[c] < 0) continue;
                                const int64_t col = (int64_t)sh.c_snr[c] * p.n_rep + sh.c_rep[c];
                                const cplx hh = cmake(hr[ct][e], hi[ct][e]);
                                if (last) sd.hdiag[col * K + i] = hh;
                                if (p.mse) {
                                    const cplx ht = p.htrue[wf][(int64_t)sh.c_rep[c] * K + i];
                                    dm[ct][e] += (hh.x - ht.x) * (hh.x - ht.x) + (hh.y - ht.y) * (hh.y - ht.y);
                                }
                                const cplx xh = cdiv_fast(yv[ct][e], hh);
                                if (!select) { vbuf[i * NC + c] = xh; continue; }
                                if (d < 0) continue;
                                const cplx xd = cmake(xh.x * sd.inv_sqrt_dpr, sd.detect_mode == 0 ? 0.0 : xh.y * sd.inv_sqrt_dpr);
                                const int word = ic_decide(cd, xd, sd.txw_t[((int64_t)(sh.c_rep[c] >> 4) * sd.n_data + d) * 16 + (sh.c_rep[c] & 15)], em, e_all[ct][e], e_edge[ct][e]);
                                if (last) sd.xD[0][col * sd.n_data + d] = xd;
                                if (next_pre) zw[ZI(d, c)] = (uint8_t)word;
                            }
                    }
                }
                if (p.mse) {
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double a = dm[ct][e];
#pragma unroll
                            for (int o = 4; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                            if (g == 0) atomicAdd(&sh.mse[ct * 8 + 2 * t4 + e], a);
                        }
                }
                if (select) {                                  // lanes with the same t4 hold the same four columns
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            unsigned a = e_all[ct][e], b = e_edge[ct][e];
#pragma unroll
                            for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
                            if (g == 0) {
                                if (a) atomicAdd(&sh.cnt[ct * 8 + 2 * t4 + e][0], a);
                                if (b) atomicAdd(&sh.cnt[ct * 8 + 2 * t4 + e][1], b);
                            }
                        }
                }
            } else if (!detected) {
                const int c = tid % NC;
                const bool okc = sh.c_rep[c] >= 0;
                const SchemeDev& sd = p.sch[sh.c_scheme[c]];
                const ConstDev& cd = sh.cst[sd.constellation];
                const bool select = sd.detect_mode != 1;
                const cplx* __restrict__ ht = p.htrue[wf] + (int64_t)(okc ? sh.c_rep[c] : 0) * K;
                const int64_t colbase = ((int64_t)sh.c_snr[c] * p.n_rep + (okc ? sh.c_rep[c] : 0)) * sd.n_data;
                const uint32_t* __restrict__ txw = sd.txword + (int64_t)(okc ? sh.c_rep[c] : 0) * sd.n_data;
                unsigned e_all = 0, e_edge = 0;
                const int istep = nthr / NC;
                for (int i0 = tid / NC; i0 < K; i0 += IC_LIGHT_UNROLL * istep) {
                    cplx yv[IC_LIGHT_UNROLL], hv[IC_LIGHT_UNROLL];
                    int dd[IC_LIGHT_UNROLL];
#pragma unroll
                    for (int u = 0; u < IC_LIGHT_UNROLL; ++u) {
                        const int i = i0 + u * istep;
                        yv[u] = (okc && i < K) ? yic(i, c) : cmake(0.0, 0.0);
                        hv[u] = (okc && i < K) ? ht[i] : cmake(1.0, 0.0);
                        dd[u] = (okc && select && i < K) ? sd.pos2data[i] : -1;
                    }
#pragma unroll
                    for (int u = 0; u < IC_LIGHT_UNROLL; ++u) {
                        const int i = i0 + u * istep;
                        if (i >= K || !okc) continue;
                        const cplx xh = cdiv_fast(yv[u], hv[u]);
                        if (!select) { vbuf[i * NC + c] = xh; continue; }
                        const int d = dd[u];
                        if (d < 0) continue;
                        const cplx xd = cmake(xh.x * sd.inv_sqrt_dpr, sd.detect_mode == 0 ? 0.0 : xh.y * sd.inv_sqrt_dpr);
                        const int word = ic_decide(cd, xd, txw[d], sd.edge_mask[d], e_all, e_edge);
                        if (last) sd.xD[1][colbase + d] = xd;
                        if (next_pre) zw[ZI(d, c)] = (uint8_t)word;
                    }
                }
                if (e_all) atomicAdd(&sh.cnt[c][0], e_all);
                if (e_edge) atomicAdd(&sh.cnt[c][1], e_edge);
            }
            __syncthreads();
        }
        // ---- phase E for de-spread schemes: x_d = C_d^H x_hat / dpr over the spreading set   (DS.m:430-433 ...)
        if (any_despread) {
            // the block size is a multiple of 16, so a thread keeps its column: counters stay in registers
            const int c = tid % NC;
            unsigned e_all = 0, e_edge = 0;
            const SchemeDev& sd = p.sch[sh.c_scheme[c]];
            if (sh.c_rep[c] >= 0 && sd.detect_mode == 1) {
                const ConstDev& cd = sh.cst[sd.constellation];
                const int64_t colbase = ((int64_t)sh.c_snr[c] * p.n_rep + sh.c_rep[c]) * sd.n_data;
                const uint32_t* __restrict__ txw = sd.txw_t + (int64_t)(sh.c_rep[c] >> 4) * sd.n_data * 16 + (sh.c_rep[c] & 15);
                const int dstep = nthr / NC;
                for (int d0 = tid / NC; d0 < sd.n_data; d0 += 4 * dstep) {
                    cplx xd[4];
                    int e0[4], e1[4];
                    cplx cv0[4], x0[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                      // four data symbols in flight per thread
                        const int d = d0 + u * dstep;
                        e0[u] = e1[u] = 0;
                        if (d < sd.n_data) { e0[u] = sd.ct_colptr[sd.P + d]; e1[u] = sd.ct_colptr[sd.P + d + 1]; }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                      // first entry of each spreading set (most have one)
                        const bool in = e0[u] < e1[u];
                        cv0[u] = in ? sd.ct_val[e0[u]] : cmake(0.0, 0.0);
                        x0[u] = in ? vbuf[sd.ct_row[e0[u]] * NC + c] : cmake(0.0, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        cplx acc = cmulc(cv0[u], x0[u]);
                        for (int e = e0[u] + 1; e < e1[u]; ++e) {
                            const cplx t = cmulc(sd.ct_val[e], vbuf[sd.ct_row[e] * NC + c]);
                            acc.x += t.x; acc.y += t.y;
                        }
                        xd[u] = cmake(acc.x * sd.inv_dpr, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int d = d0 + u * dstep;
                        if (d < sd.n_data) {
                            const int word = ic_decide(cd, xd[u], txw[(int64_t)d * 16], sd.edge_mask[d], e_all, e_edge);
                            if (last) sd.xD[csi][colbase + d] = xd[u];
                            if (next_pre) zw[ZI(d, c)] = (uint8_t)word;
                        }
                    }
                }
            }
            if (e_all) atomicAdd(&sh.cnt[c][0], e_all);
            if (e_edge) atomicAdd(&sh.cnt[c][1], e_edge);
            __syncthreads();                                   // readers of vbuf done, decided words complete
        }
        if (tid < 2 * NC && !detected) {
            int cc = tid >> 1, e = tid & 1;
            if (cc < cta.n_cols && sh.c_rep[cc] >= 0) {
                int64_t o = ((((int64_t)sh.c_rep[cc] * p.n_snr + sh.c_snr[cc]) * (p.n_iter + 1) + it) * 3 + sh.c_scheme[cc]) * 4 + csi * 2 + e;
                p.err[o] = sh.cnt[cc][e];
            }
        }
        if (p.mse && csi == 0 && tid < cta.n_cols && sh.c_rep[tid] >= 0)
            p.mse[(((int64_t)sh.c_rep[tid] * p.n_snr + sh.c_snr[tid]) * (p.n_iter + 1) + it) * 3 + sh.c_scheme[tid]] = sh.mse[tid];
        if (!next_pre) continue;
        // ---- phase A of iteration it+1: v = C z with z = [xP; decided symbols]   (DS.m:482-484, 541-543)
        if (cm) {
            // column-major units: a thread owns a row and walks the columns, so the row's precoder entry is read once per scheme
            // and every store of a warp is one contiguous run of a column; the next row's entry is loaded under the current stores
            const int nh = cta.n_cols > 8 ? 2 : 1;             // an 8-column half is on one scheme
            const SchemeDev& sA = p.sch[sh.c_scheme[0]];
            const SchemeDev& sB = p.sch[sh.c_scheme[nh > 1 ? 8 : 0]];
            const bool same = nh < 2 || sh.c_scheme[8] == sh.c_scheme[0];
            int cA = -2, cB = -2; cplx vA = cmake(0.0, 0.0), vB = cmake(0.0, 0.0);
            if (tid < K) { cA = sA.row_col0[tid]; vA = sA.row_val0[tid]; if (!same) { cB = sB.row_col0[tid]; vB = sB.row_val0[tid]; } }
            for (int i = tid; i < K; i += nthr) {
                const int in = i + nthr;
                int nA = -2, nB = -2; cplx wA = cmake(0.0, 0.0), wB = cmake(0.0, 0.0);
                if (in < K) { nA = sA.row_col0[in]; wA = sA.row_val0[in]; if (!same) { nB = sB.row_col0[in]; wB = sB.row_val0[in]; } }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half >= nh) break;
                    const SchemeDev& sd = half ? sB : sA;
                    const int col0 = (half && !same) ? cB : cA;
                    const cplx val0 = (half && !same) ? vB : vA;
                    if (col0 == -2) continue;
                    const cplx* sym = sh.cst[sd.constellation].symbol;
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        const int c = half * 8 + cc;
                        if (sh.c_rep[c] < 0) continue;
                        const cplx zz = col0 < 0 ? cmake(0.0, 0.0) : (col0 < sd.P ? xPs[col0 * NC + c] : sym[zw[ZI(col0 - sd.P, c)]]);
                        vbuf[VI(i, c)] = cmul(val0, zz);
                    }
                }
                cA = nA; vA = wA; cB = nB; vB = wB;
            }
        } else {   // rows with at most one entry: v[i] = val0[i] * z[col0[i]]; eight rows in flight per thread
            const int istep = nthr / NC;
            const int c = tid % NC;
            const bool okc = sh.c_rep[c] >= 0;
            const SchemeDev& sd = p.sch[sh.c_scheme[c]];
            const cplx* sym = sh.cst[sd.constellation].symbol;
            for (int i0 = tid / NC; i0 < K; i0 += 8 * istep) {
                int col[8]; cplx val[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * istep;
                    col[u] = (okc && i < K) ? sd.row_col0[i] : -1;
                    val[u] = (okc && i < K) ? sd.row_val0[i] : cmake(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * istep;
                    if (i < K && col[u] != -2) {
                        const cplx zz = col[u] < 0 ? cmake(0.0, 0.0)
                                       : (col[u] < sd.P ? xPs[col[u] * NC + c] : sym[zw[ZI(col[u] - sd.P, c)]]);
                        vbuf[VI(i, c)] = cmul(val[u], zz);
                    }
                }
            }
        }
        // long rows (auxiliary symbols, spread symbols) on the FP64 tensor pipe: a warp takes a tile of 8 long
        // rows; per k-step the A fragment holds their coefficients for four z rows (the tile's column union,
        // packed by the host), the B fragment the z values of those rows for the 8 columns of an n-tile.
        // An n-tile (8 columns) is always on one scheme; the two halves of an EST unit share the A fragments.
        {
            const int g = lane >> 2, t4 = lane & 3;
            const int s0 = sh.c_scheme[0], s1 = cta.n_cols > 8 ? sh.c_scheme[8] : -1;
            for (int pass = 0; pass < 2; ++pass) {
                // pass 0: the scheme of half 0, on both halves if they agree; pass 1: half 1 when its scheme differs
                if (pass == 1 && (s1 < 0 || s1 == s0)) break;
                const int si = pass == 0 ? s0 : s1;
                const int ct0 = pass, ct1 = (pass == 0 && s1 == s0) ? 2 : pass + 1;
                const SchemeDev& sd = p.sch[si];
                const cplx* sym = sh.cst[sd.constellation].symbol;
                const int P = sd.P;
                for (int tl = warp; tl < sd.n_lr_tiles; tl += nwarp) {
                    double cr[2][2] = {{0, 0}, {0, 0}}, ci[2][2] = {{0, 0}, {0, 0}};
                    const int st1 = sd.lr_ptr[tl + 1];
                    for (int st = sd.lr_ptr[tl]; st < st1; ++st) {
                        const cplx a = ld_stream(sd.lr_frag + (int64_t)st * 32 + lane);
                        const int kc = sd.lr_kcol[st * 4 + t4];
                        const double nai = dneg(a.y);
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct) {
                            if (ct < ct0 || ct >= ct1) continue;
                            const int c = ct * 8 + g;
                            const cplx b = kc < P ? xPs[kc * NC + c] : sym[zw[ZI(kc - P, c)]];
                            dmma884(cr[ct][0], cr[ct][1], a.x, b.x);
                            dmma884(cr[ct][0], cr[ct][1], nai, b.y);
                            dmma884(ci[ct][0], ci[ct][1], a.x, b.y);
                            dmma884(ci[ct][0], ci[ct][1], a.y, b.x);
                        }
                    }
                    const int r = tl * 8 + g;
                    if (r < sd.n_long_rows) {
                        const int i = sd.long_rows[r];
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = ct * 8 + 2 * t4 + e;
                                if (ct >= ct0 && ct < ct1 && sh.c_rep[c] >= 0) vbuf[VI(i, c)] = cmake(cr[ct][e], ci[ct][e]);
                            }
                    }
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// k_ic_post: phases C, D, E of iteration `it` and phase A of iteration it+1 in (mostly) ONE pass over the unit's
// K x 16 tile of cancelled symbols (the successor of k_ic_light, which stays selectable: CHEST_LIGHT=old).
//   * y_ic is streamed through a shared-memory ring of 64-row chunks (16 KB) filled by bulk asynchronous copies
//     (cp.async.bulk + mbarrier complete_tx: one elected thread, LOOK chunks ahead); a warp copies its row tile from
//     the ring into registers and frees the stage at once, so the memory-level parallelism does not depend on registers;
//   * the operands a row tile needs besides y (diagonal-W fragments, row descriptors, true channel) are prefetched into
//     registers one row tile ahead;
//   * a row is equalised, decided and -- for rows whose precoder entry refers to the symbol decided at that very row
//     (data rows of all three schemes, pilot rows) -- precoded for the next iteration right away: one read of y_ic, one
//     write of v.  Only the spread symbols of the data-spreading scheme (gather over their spreading sets) and the long
//     precoder rows (auxiliary / spread positions, DMMA tiles) need a second look, after one block barrier.
#ifndef POST_NS
#define POST_NS 4              // ring stages
#endif
#ifndef POST_LOOK
#define POST_LOOK 2            // chunks in flight ahead of the consumers
#endif
#define POST_ROWS 64           // rows per chunk = 8 warps x 8 rows
#define POST_THREADS 256
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }"
                 :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                    "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int P4T>
__global__ void __launch_bounds__(POST_THREADS, 2) k_ic_post(IcParams p) {
    constexpr int NC = NC_MAX, HS = NC + 2, NS = POST_NS, LOOK = POST_LOOK, CH = POST_ROWS * NC;
    constexpr int PQ = P4T > 0 ? P4T : 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int it = p.it;
    extern __shared__ __align__(128) cplx ic_smem[];
    cplx* ring = ic_smem;                                   // [NS][64 rows][16 columns]
    cplx* hPn = ring + NS * CH;                             // new pilot estimates [p][col] (zero padded to 4*P4 rows)
    cplx* xPs = hPn + p.pilot_rows * HS;                    // transmitted pilots of the unit's columns [p][col]
    __shared__ IcShared sh;
    __shared__ uint64_t full[NS], empty[NS];
    uint8_t* zw;                                            // decided words of the data symbols [d][col]
    {
        cplx* q = xPs + p.pilot_rows * NC;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const ConstDev& cg = p.cst[k];
            cplx* sym = q; q += cg.order;
            double* lev = reinterpret_cast<double*>(q); q += (cg.n_axis + 1) / 2;
            int* gr = reinterpret_cast<int*>(q); q += (cg.order + 3) / 4;
            for (int e = tid; e < cg.order; e += POST_THREADS) { sym[e] = cg.symbol[e]; gr[e] = cg.word_of_grid[e]; }
            for (int e = tid; e < cg.n_axis; e += POST_THREADS) lev[e] = cg.level[e];
            if (tid == 0) { sh.cst[k] = cg; sh.cst[k].symbol = sym; sh.cst[k].level = lev; sh.cst[k].word_of_grid = gr; }
        }
        zw = reinterpret_cast<uint8_t*>(q);
    }
    if (tid == 0) {
        for (int s_ = 0; s_ < NS; ++s_) { mbar_init(&full[s_], 1); mbar_init(&empty[s_], POST_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool last = it == p.n_iter, next_pre = !last;
    unsigned seq = 0;                                       // chunks this CTA has streamed so far (ring position)
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const IcCta cta = p.ctas[unit];
        const int csi = cta.mode;
        const int wf = (cta.mode == 0) ? p.sch[cta.scheme_or_wf].waveform : cta.scheme_or_wf;
        const int K = p.sch[p.wf_scheme[wf][0]].K, RT = (K + 7) / 8;
        cplx* vbuf = p.scratch + (int64_t)unit * 3 * p.K_max * NC + (int64_t)p.K_max * NC;
        const cplx* ybuf = vbuf + (int64_t)p.K_max * NC;
        const int ngroup = (RT + 7) / 8;                    // groups of 8 row tiles = chunks of 64 rows
        const bool staged = it > 0;
        auto issue = [&](int j) {                           // chunk j of this unit -> ring (thread 0 only)
            const unsigned q = seq + j, st = q % NS;
            if (q >= NS) mbar_wait(&empty[st], ((q / NS) - 1) & 1);
            const unsigned rows = min(POST_ROWS, K - j * POST_ROWS);
            mbar_expect_tx(&full[st], rows * NC * 16);
            bulk_g2s(ring + st * CH, ybuf + (int64_t)j * CH, rows * NC * 16, &full[st]);
        };
        if (staged && tid == 0)
            for (int j = 0; j < LOOK && j < ngroup; ++j) issue(j);
        ic_load_unit(p, cta, sh);
        const int Pw = p.sch[p.wf_scheme[wf][0]].P;
        // ---- transmitted pilots of the columns; phase C: LS pilot estimates (DS.m:412-414, 487-489)
        for (int idx = tid; idx < p.pilot_rows * NC; idx += POST_THREADS) {
            const int c = idx % NC, pp = idx / NC;
            const bool ok = sh.c_rep[c] >= 0 && pp < Pw;
            const cplx xp = ok ? p.sch[sh.c_scheme[c]].xP[(int64_t)sh.c_rep[c] * Pw + pp] : cmake(1.0, 0.0);
            xPs[pp * NC + c] = xp;
            if (csi == 0) {
                const SchemeDev& sd = p.sch[cta.scheme_or_wf];
                cplx hp = cmake(0.0, 0.0);
                if (ok) {
                    const int i = sd.pilot_pos[pp];
                    const cplx yv = staged ? ybuf[i * NC + c] : sh.ycolp[c][i];
                    const cplx q = cdiv(yv, xp);
                    hp = cmake(q.x / sd.sqrt_kappa, q.y / sd.sqrt_kappa);
                    sd.hP[((int64_t)sh.c_snr[c] * p.n_rep + sh.c_rep[c]) * sd.P + pp] = hp;
                }
                hPn[pp * HS + c] = hp;
            }
        }
        __syncthreads();
        // ---- the pass: equalise, decide, count, precode
        const int var_cur = (it == 0 || it <= p.n_iter / 2) ? 0 : 1;
        const int sc0 = sh.c_scheme[0], sc1 = cta.n_cols > 8 ? sh.c_scheme[8] : sc0;
        const SchemeDev& sd0 = p.sch[sc0];
        const SchemeDev& sd1 = p.sch[sc1];
        const bool two = sc1 != sc0;                        // PERF unit with two schemes: one per 8-column half
        const int P4 = sd0.P4;
        const cplx* __restrict__ wfr = csi == 0 ? sd0.wdiag_frag[var_cur] + (int64_t)cta.snr * RT * P4 * 32 + lane : nullptr;
        const cplx* __restrict__ ht = csi == 1 ? p.htrue[wf] + (int64_t)cta.snr * K : nullptr;     // cta.snr holds the realization
        struct TileIn { cplx w[PQ]; int4 ri[2]; cplx val0[2]; cplx h; };
        auto load_in = [&](int rt, TileIn& in) {
            const int i = min(rt * 8 + g, K - 1);
            if (csi == 0) {
                if (P4T > 0) {
#pragma unroll
                    for (int pq = 0; pq < PQ; ++pq) in.w[pq] = ld_stream(wfr + ((int64_t)rt * P4T + pq) * 32);
                }
            } else in.h = ht[i];
            in.ri[0] = sd0.rowinfo[i]; in.val0[0] = sd0.row_val0[i];
            if (two) { in.ri[1] = sd1.rowinfo[i]; in.val0[1] = sd1.row_val0[i]; }
            else { in.ri[1] = in.ri[0]; in.val0[1] = in.val0[0]; }
        };
        unsigned e_all[2][2] = {{0, 0}, {0, 0}}, e_edge[2][2] = {{0, 0}, {0, 0}};
        TileIn nxt;
        if (warp < RT) load_in(warp, nxt);
#pragma unroll 1
        for (int j = 0; j < ngroup; ++j) {
            const int rt = j * 8 + warp;
            const bool have = rt < RT;
            TileIn cur = nxt;
            if (rt + 8 < RT) load_in(rt + 8, nxt);
            cplx y[2][2];
            const int i = rt * 8 + g;
            if (staged) {
                if (tid == 0 && j + LOOK < ngroup) issue(j + LOOK);
                const unsigned q = seq + j, st = q % NS;
                mbar_wait(&full[st], (q / NS) & 1);
                if (have) {
                    const cplx* yr = ring + st * CH + (warp * 8 + g) * NC + 2 * t4;
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct) { y[ct][0] = yr[ct * 8]; y[ct][1] = yr[ct * 8 + 1]; }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);     // the values are in registers: the stage is free
            } else if (have) {
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const cplx* yp = sh.ycolp[ct * 8 + 2 * t4 + e];
                        y[ct][e] = (yp && i < K) ? yp[i] : cmake(1.0, 0.0);
                    }
            }
            if (!have) continue;
            const bool rowok = i < K;
            // one-tap channel of the row: h_est = W_diag * hP as a small DMMA product, or the true channel
            cplx hh[2][2];
            if (csi == 0) {
                double hr[2][2] = {{0, 0}, {0, 0}}, hi[2][2] = {{0, 0}, {0, 0}};
                const int npq = P4T > 0 ? P4T : P4;
#pragma unroll
                for (int pq = 0; pq < npq; ++pq) {
                    const cplx a = P4T > 0 ? cur.w[P4T > 0 ? pq : 0] : ld_stream(wfr + ((int64_t)rt * P4 + pq) * 32);
                    const double nai = dneg(a.y);
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct) {
                        const cplx b = hPn[(pq * 4 + t4) * HS + ct * 8 + g];
                        dmma884(hr[ct][0], hr[ct][1], a.x, b.x);
                        dmma884(hr[ct][0], hr[ct][1], nai, b.y);
                        dmma884(hi[ct][0], hi[ct][1], a.x, b.y);
                        dmma884(hi[ct][0], hi[ct][1], a.y, b.x);
                    }
                }
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) hh[ct][e] = cmake(hr[ct][e], hi[ct][e]);
            } else {
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) hh[ct][0] = hh[ct][1] = cur.h;
            }
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                const SchemeDev& sd = ct == 0 ? sd0 : sd1;
                const ConstDev& cd = sh.cst[sd.constellation];
                const int4 ri = cur.ri[ct];
                const cplx val0 = cur.val0[ct];
                const int mode = sd.detect_mode;
                cplx vv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = ct * 8 + 2 * t4 + e;
                    const bool valid = rowok && sh.c_rep[c] >= 0;
                    const cplx xh = cdiv(y[ct][e], hh[ct][e]);
                    int word = 0;
                    vv[e] = cmake(0.0, 0.0);
                    if (valid) {
                        const int64_t col = (int64_t)sh.c_snr[c] * p.n_rep + sh.c_rep[c];
                        if (last && csi == 0) sd.hdiag[col * K + i] = hh[ct][e];
                        if (ri.x >= 0) {
                            const int d = ri.x;
                            const cplx xd = mode == 1 ? cmake((val0.x * xh.x + val0.y * xh.y) / sd.dpr, 0.0)
                                                      : cmake(xh.x / sd.sqrt_dpr, mode == 0 ? 0.0 : xh.y / sd.sqrt_dpr);
                            const uint32_t tw = sd.txw_t[((int64_t)(sh.c_rep[c] >> 4) * sd.n_data + d) * 16 + (sh.c_rep[c] & 15)];
                            word = ic_decide(cd, xd, tw, (uint32_t)ri.z, e_all[ct][e], e_edge[ct][e]);
                            if (last) sd.xD[csi][col * sd.n_data + d] = xd;
                            if (next_pre) zw[d * NC + c] = (uint8_t)word;
                        }
                        if (ri.w & 1) vbuf[i * NC + c] = xh;            // wanted by the de-spreading gather below
                    }
                    if (next_pre && sd.fuse_ok && ri.y != -2) {         // v = C z for rows with at most one entry
                        const cplx zz = ri.y < 0 ? cmake(0.0, 0.0) : (ri.y < sd.P ? xPs[ri.y * NC + c] : cd.symbol[word]);
                        vv[e] = cmul(val0, zz);
                    }
                }
                if (next_pre && sd.fuse_ok && ri.y != -2 && rowok && !(ri.w & 1)) st_cplx2(vbuf + i * NC + ct * 8 + 2 * t4, vv[0], vv[1]);
            }
        }
        seq += staged ? ngroup : 0;
        // lanes with the same t4 hold the same four columns
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                unsigned a = e_all[ct][e], b = e_edge[ct][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
                if (g == 0) {
                    if (a) atomicAdd(&sh.cnt[ct * 8 + 2 * t4 + e][0], a);
                    if (b) atomicAdd(&sh.cnt[ct * 8 + 2 * t4 + e][1], b);
                }
            }
        __syncthreads();
        // ---- data symbols spread over several positions: x_d = C_d^H x_hat / dpr over the spreading set (DS.m:436-437, 520)
        if ((sd0.detect_mode == 1 && sd0.n_multi > 0) || (sd1.detect_mode == 1 && sd1.n_multi > 0)) {
            const int c = tid % NC;
            unsigned ea = 0, ee = 0;
            const SchemeDev& sd = p.sch[sh.c_scheme[c]];
            if (sh.c_rep[c] >= 0 && sd.detect_mode == 1) {
                const ConstDev& cd = sh.cst[sd.constellation];
                const int64_t colbase = ((int64_t)sh.c_snr[c] * p.n_rep + sh.c_rep[c]) * sd.n_data;
                const uint32_t* __restrict__ txw = sd.txw_t + (int64_t)(sh.c_rep[c] >> 4) * sd.n_data * 16 + (sh.c_rep[c] & 15);
                constexpr int DSTEP = POST_THREADS / NC;
                for (int k0 = tid / NC; k0 < sd.n_multi; k0 += 4 * DSTEP) {
                    int dd[4], e0[4], e1[4];
                    cplx acc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + u * DSTEP;
                        dd[u] = k < sd.n_multi ? sd.multi_d[k] : -1;
                        e0[u] = e1[u] = 0;
                        if (dd[u] >= 0) { e0[u] = sd.ct_colptr[sd.P + dd[u]]; e1[u] = sd.ct_colptr[sd.P + dd[u] + 1]; }
                        acc[u] = cmake(0.0, 0.0);
                    }
                    int nmax = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) nmax = max(nmax, e1[u] - e0[u]);
                    for (int q = 0; q < nmax; ++q) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (e0[u] + q < e1[u]) {
                                const cplx t = cmulc(sd.ct_val[e0[u] + q], vbuf[sd.ct_row[e0[u] + q] * NC + c]);
                                acc[u].x += t.x; acc[u].y += t.y;
                            }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (dd[u] >= 0) {
                            const cplx xd = cmake(acc[u].x / sd.dpr, 0.0);
                            const int word = ic_decide(cd, xd, txw[(int64_t)dd[u] * 16], sd.edge_mask[dd[u]], ea, ee);
                            if (last) sd.xD[csi][colbase + dd[u]] = xd;
                            if (next_pre) zw[dd[u] * NC + c] = (uint8_t)word;
                        }
                }
            }
            if (ea) atomicAdd(&sh.cnt[c][0], ea);
            if (ee) atomicAdd(&sh.cnt[c][1], ee);
            __syncthreads();                                   // readers of vbuf done, decided words complete
        }
        if (tid < 2 * NC) {
            const int cc = tid >> 1, e = tid & 1;
            if (cc < cta.n_cols && sh.c_rep[cc] >= 0) {
                const int64_t o = ((((int64_t)sh.c_rep[cc] * p.n_snr + sh.c_snr[cc]) * (p.n_iter + 1) + it) * 3 + sh.c_scheme[cc]) * 4 + csi * 2 + e;
                p.err[o] = sh.cnt[cc][e];
            }
        }
        if (!next_pre) continue;
        // ---- phase A leftovers: single-entry rows of schemes whose precoder does not allow the fused write ...
        if (!sd0.fuse_ok || !sd1.fuse_ok) {
            const int c = tid % NC;
            const SchemeDev& sd = p.sch[sh.c_scheme[c]];
            if (sh.c_rep[c] >= 0 && !sd.fuse_ok) {
                const cplx* sym = sh.cst[sd.constellation].symbol;
                for (int i = tid / NC; i < K; i += POST_THREADS / NC) {
                    const int col = sd.row_col0[i];
                    if (col == -2) continue;
                    const cplx zz = col < 0 ? cmake(0.0, 0.0) : (col < sd.P ? xPs[col * NC + c] : sym[zw[(col - sd.P) * NC + c]]);
                    vbuf[i * NC + c] = cmul(sd.row_val0[i], zz);
                }
            }
        }
        // ... and the long rows (auxiliary symbols, spread positions) on the FP64 tensor pipe: a warp takes a tile of 8 long
        // rows; per k-step the A fragment holds their coefficients for four z rows (the tile's column union, packed by
        // the host), the B fragment the z values of those rows for the 8 columns of an n-tile.
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1 && !two) break;
            const SchemeDev& sd = pass == 0 ? sd0 : sd1;
            const int ct0 = pass, ct1 = (pass == 0 && !two) ? 2 : pass + 1;
            const cplx* sym = sh.cst[sd.constellation].symbol;
            const int P = sd.P;
            for (int tl = warp; tl < sd.n_lr_tiles; tl += POST_THREADS / 32) {
                double cr[2][2] = {{0, 0}, {0, 0}}, ci[2][2] = {{0, 0}, {0, 0}};
                const int st1 = sd.lr_ptr[tl + 1];
                for (int st = sd.lr_ptr[tl]; st < st1; ++st) {
                    const cplx a = ld_stream(sd.lr_frag + (int64_t)st * 32 + lane);
                    const int kc = sd.lr_kcol[st * 4 + t4];
                    const double nai = dneg(a.y);
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct) {
                        if (ct < ct0 || ct >= ct1) continue;
                        const int c = ct * 8 + g;
                        const cplx b = kc < P ? xPs[kc * NC + c] : sym[zw[(kc - P) * NC + c]];
                        dmma884(cr[ct][0], cr[ct][1], a.x, b.x);
                        dmma884(cr[ct][0], cr[ct][1], nai, b.y);
                        dmma884(ci[ct][0], ci[ct][1], a.x, b.y);
                        dmma884(ci[ct][0], ci[ct][1], a.y, b.x);
                    }
                }
                const int r = tl * 8 + g;
                if (r < sd.n_long_rows) {
                    const int i = sd.long_rows[r];
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct)
                        if (ct >= ct0 && ct < ct1) st_cplx2(vbuf + i * NC + ct * 8 + 2 * t4, cmake(cr[ct][0], ci[ct][0]), cmake(cr[ct][1], ci[ct][1]));
                }
            }
        }
    }
}


// ============================================================================ FFT modem (FBMC.m:255-315, OFDM.m:153-181)
// Hand-written mixed-radix Stockham FFT in shared memory (sizes with prime factors <= 13: 24, 168, 196, 512, ... ; no
// cuFFT), one thread block per (multicarrier symbol, column).  tw[m] = exp(-2 pi i m / n); the inverse transform
// conjugates the table and scales by 1/n like MATLAB's ifft.
#define FFT_MAX_STAGES 16
#ifndef PERF_FBMC_THREADS
#define PERF_FBMC_THREADS 128
#endif
#ifndef PERF_FBMC_MIN_CTAS
#define PERF_FBMC_MIN_CTAS 8
#endif
struct FftPlan { int n, n_stage; int radix[FFT_MAX_STAGES]; };
struct ModemDev {
    int kind;                    // 0: FBMC polyphase, 1: CP-OFDM
    int L, Ksym, nfft, N, time_spacing, O, cp, zero_guard, Np;
    double norm, inv_demod;      // NormalizationFactor; 1 / (NormalizationFactor * SubcarrierSpacing) (FBMC) or 1 / NormalizationFactor (OFDM)
    const int* bin;              // FFT bin of subcarrier l (IndexPolyphaseMap rows in ascending order, FBMC.m:154-156)
    const double* filt;          // prototype filter, Np = O * nfft real taps (FBMC)
    const cplx* phase;           // PhaseShift, L x Ksym (FBMC.m:139)
    const cplx* tw;              // exp(-2 pi i m / nfft), m = 0..nfft-1
    FftPlan plan;
};
template <int R>
__device__ __forceinline__ void fft_butterfly(const cplx* __restrict__ a, cplx* __restrict__ b, const cplx* __restrict__ tw,
                                              int n, int Ns, int j, int m, bool inv) {
    const int k = j % Ns, step = n / (Ns * R);
    cplx v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
        cplx w = tw[(int)(((int64_t)t * k * step) % n)];
        if (inv) w.y = -w.y;
        v[t] = cmul(a[j + t * m], w);
    }
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int u = 0; u < R; ++u) {
        cplx acc = v[0];
#pragma unroll
        for (int t = 1; t < R; ++t) {
            cplx w = tw[((t * u) % R) * (n / R)];
            if (inv) w.y = -w.y;
            cfma(acc, v[t], w);
        }
        b[j0 + u * Ns] = acc;
    }
}
// In-place interface: data in `a`, scratch `b` (both n complex, shared memory).  Returns the buffer holding the result.
__device__ cplx* fft_shared(cplx* a, cplx* b, const cplx* tw, const FftPlan& plan, bool inv) {
    const int n = plan.n;
    int Ns = 1;
    for (int s = 0; s < plan.n_stage; ++s) {
        const int R = plan.radix[s], m = n / R;
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            switch (R) {
                case 2: fft_butterfly<2>(a, b, tw, n, Ns, j, m, inv); break;
                case 3: fft_butterfly<3>(a, b, tw, n, Ns, j, m, inv); break;
                case 4: fft_butterfly<4>(a, b, tw, n, Ns, j, m, inv); break;
                case 5: fft_butterfly<5>(a, b, tw, n, Ns, j, m, inv); break;
                case 7: fft_butterfly<7>(a, b, tw, n, Ns, j, m, inv); break;
                case 11: fft_butterfly<11>(a, b, tw, n, Ns, j, m, inv); break;
                default: fft_butterfly<13>(a, b, tw, n, Ns, j, m, inv); break;
            }
        }
        __syncthreads();
        cplx* t_ = a; a = b; b = t_;
        Ns *= R;
    }
    return a;
}
// x: [col][L*Ksym] symbols (column-major L x Ksym per column).  FBMC: Z[col][k][nfft] = ifft of symbol k (FBMC.m:263-267
// before the tiling); OFDM: the time signal with cyclic prefix written straight into s (OFDM.m:158-164).
__global__ void k_modem_ifft(ModemDev md, const cplx* __restrict__ x, cplx* __restrict__ Z, cplx* __restrict__ s) {
    extern __shared__ __align__(16) cplx fsm[];
    cplx* a = fsm; cplx* b = fsm + md.nfft; cplx* tw = fsm + 2 * md.nfft;
    const int k = blockIdx.x, col = blockIdx.y, n = md.nfft;
    for (int m = threadIdx.x; m < n; m += blockDim.x) { a[m] = cmake(0.0, 0.0); tw[m] = md.tw[m]; }
    __syncthreads();
    for (int l = threadIdx.x; l < md.L; l += blockDim.x) {
        cplx v = x[((int64_t)col * md.Ksym + k) * md.L + l];
        if (md.kind == 0) v = cmul(v, md.phase[k * md.L + l]);
        a[md.bin[l]] = cmake(v.x * md.norm, v.y * md.norm);
    }
    __syncthreads();
    const cplx* z = fft_shared(a, b, tw, md.plan, true);
    const double sc = 1.0 / n;
    if (md.kind == 0) {
        for (int m = threadIdx.x; m < n; m += blockDim.x) Z[((int64_t)col * md.Ksym + k) * n + m] = cmake(z[m].x * sc, z[m].y * sc);
    } else {
        const int T = md.time_spacing;                          // FFT size + cyclic prefix
        cplx* out = s + (int64_t)col * md.N + md.zero_guard + (int64_t)k * T;
        for (int j = threadIdx.x; j < T; j += blockDim.x) {
            const int m = (j - md.cp + n) % n;
            out[j] = cmake(z[m].x * sc, z[m].y * sc);
        }
    }
}
// FBMC overlap-add as a gather: s[n] = sum_k p[n - k T] * Z[k][(n - k T) mod nfft], symbols ascending (FBMC.m:267-268)
__global__ void k_fbmc_overlap_add(ModemDev md, const cplx* __restrict__ Z, cplx* __restrict__ s) {
    const int n = blockIdx.y * blockDim.x + threadIdx.x, col = blockIdx.x;      // columns on grid.x: up to 2^31 - 1 of them
    if (n >= md.N) return;
    const int T = md.time_spacing;
    int k_lo = (n - md.Np + T) / T; if (n - md.Np + 1 <= 0) k_lo = 0;      // smallest k with n - kT <= Np - 1
    const int k_hi = min(md.Ksym - 1, n / T);
    cplx acc = cmake(0.0, 0.0);
    for (int k = k_lo; k <= k_hi; ++k) {
        const int tap = n - k * T;
        if (tap < 0 || tap >= md.Np) continue;
        const cplx z = Z[((int64_t)col * md.Ksym + k) * md.nfft + tap % md.nfft];
        const double pf = md.filt[tap];
        acc.x += pf * z.x; acc.y += pf * z.y;
    }
    s[(int64_t)col * md.N + n] = acc;
}
// Demodulation: FBMC: filter, fold by O, FFT, pick the subcarrier bins, undo the phase (FBMC.m:294-302);
// OFDM: drop the cyclic prefix, FFT, pick (OFDM.m:172-180).  y: [col][L*Ksym].
__global__ void k_modem_fft(ModemDev md, const cplx* __restrict__ r, cplx* __restrict__ y) {
    extern __shared__ __align__(16) cplx fsm[];
    cplx* a = fsm; cplx* b = fsm + md.nfft; cplx* tw = fsm + 2 * md.nfft;
    const int k = blockIdx.x, col = blockIdx.y, n = md.nfft;
    const cplx* rc = r + (int64_t)col * md.N;
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        tw[m] = md.tw[m];
        cplx acc = cmake(0.0, 0.0);
        if (md.kind == 0) {
            const cplx* seg = rc + (int64_t)k * md.time_spacing;
            for (int o = 0; o < md.O; ++o) {
                const double pf = md.filt[o * n + m];
                const cplx v = seg[o * n + m];
                acc.x += v.x * pf; acc.y += v.y * pf;
            }
        } else acc = rc[md.zero_guard + (int64_t)k * md.time_spacing + md.cp + m];
        a[m] = acc;
    }
    __syncthreads();
    const cplx* z = fft_shared(a, b, tw, md.plan, false);
    for (int l = threadIdx.x; l < md.L; l += blockDim.x) {
        cplx v = z[md.bin[l]];
        if (md.kind == 0) v = cmulc(md.phase[k * md.L + l], v);        // .* conj(PhaseShift)
        y[((int64_t)col * md.Ksym + k) * md.L + l] = cmake(v.x * md.inv_demod, v.y * md.inv_demod);
    }
}


// ============================================================================ perfect-CSI pass through the polyphase modem
// y_ic = y - Q^H H G v + h v  (DS.m:541-543 with D = Q^H H G never formed) for FBMC columns, with G and Q^H applied in
// their FACTORED form -- G v = Modulation(v) (FBMC.m:255-268: phase, IFFT per symbol, prototype filter, overlap-add) and
// Q^H r = Demodulation(r) (FBMC.m:287-302: filter, fold by O, FFT, conj phase) -- instead of as support-aware GEMMs:
// about 1/13 of the flops (per column 2 x (Ksym FFTs of size nfft + 2 O nfft Ksym real-complex multiply-adds) against
// 2 x 8 Np L Ksym) and no G / Q operand traffic at all.  One CTA takes CW columns that sit side by side in one unit's
// interleaved scratch (16 CW contiguous bytes per symbol row; CW = 1 measured best: more, smaller CTAs overlap their
// phases) and keeps the whole chain
//     v -> Z = IFFT -> s (overlap-add) -> r = H s (banded, the realization's taps) -> fold -> FFT -> y_ic
// in shared memory; every step is FP64 on the scalar pipe.  Used when chest_set_modem described the waveform next to its
// dense matrices, the description reproduces G / Q (checked at finalize) and the buffers fit; otherwise k_gemm_ring.
struct PerfFbmcParams {
    ModemDev md;
    int n_groups, T, N, K;
    const int2* groups;            // (first column, number of columns <= CW): columns of one (realization, scheme slot, SNR block)
    const int64_t* voff; const int64_t* yoff; const int* rep;
    const cplx* v_base; cplx* y_base;              // unit scratch: v and, one buffer behind, y_ic
    const cplx* y; const cplx* htrue; const cplx* h; const int* tap_delay;
};
template <int R>
__device__ __forceinline__ void fft_bfly(const cplx* __restrict__ a, cplx* __restrict__ b, const cplx* __restrict__ tw,
                                         int n, int Ns, int j, int m, int step, bool inv) {
    const int k = j % Ns;
    cplx v[R];
    v[0] = a[j];
#pragma unroll
    for (int t = 1; t < R; ++t) {
        v[t] = a[j + t * m];
        if (k) {                                                // first stage (Ns = 1): every input twiddle is 1
            cplx w = tw[t * k * step];                          // t k step < n: no reduction needed
            if (inv) w.y = -w.y;
            v[t] = cmul(v[t], w);
        }
    }
    cplx* o = b + (j - k) * R + k;
    if (R == 2) {
        o[0] = cadd(v[0], v[1]); o[Ns] = csub(v[0], v[1]);
    } else if (R == 4) {                                        // roots of unity 1, -+i, -1, +-i: additions only
        const cplx s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]), s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
        const cplx r = inv ? cmake(-d13.y, d13.x) : cmake(d13.y, -d13.x);      // (+-i)^-1 ... forward: -i d13, inverse: +i d13
        o[0] = cadd(s02, s13); o[Ns] = cadd(d02, r); o[2 * Ns] = csub(s02, s13); o[3 * Ns] = csub(d02, r);
    } else if (R == 3) {                                        // w = -1/2 -+ i sqrt(3)/2
        const double h3 = 0.86602540378443864676;
        const cplx sm = cadd(v[1], v[2]), df = csub(v[1], v[2]);
        const cplx base = cmake(v[0].x - 0.5 * sm.x, v[0].y - 0.5 * sm.y);
        const cplx rot = inv ? cmake(-h3 * df.y, h3 * df.x) : cmake(h3 * df.y, -h3 * df.x);     // forward: -i h3 df
        o[0] = cadd(v[0], sm); o[Ns] = cadd(base, rot); o[2 * Ns] = csub(base, rot);
    } else {
#pragma unroll
        for (int u = 0; u < R; ++u) {
            cplx acc = v[0];
#pragma unroll
            for (int t = 1; t < R; ++t) {
                cplx w = tw[((t * u) % R) * (n / R)];
                if (inv) w.y = -w.y;
                cfma(acc, v[t], w);
            }
            o[u * Ns] = acc;
        }
    }
}
// n_batch independent transforms of size plan.n stored back to back in `a` (scratch `b`, same size); returns the result buffer
__device__ cplx* fft_shared_batch(cplx* a, cplx* b, const cplx* tw, const FftPlan& plan, bool inv, int n_batch) {
    const int n = plan.n;
    int Ns = 1;
    for (int s = 0; s < plan.n_stage; ++s) {
        const int R = plan.radix[s], m = n / R, step = n / (Ns * R);
        for (int idx = threadIdx.x; idx < n_batch * m; idx += blockDim.x) {
            const int f = idx / m, j = idx - f * m;
            const cplx* af = a + f * n; cplx* bf = b + f * n;
            switch (R) {
                case 2: fft_bfly<2>(af, bf, tw, n, Ns, j, m, step, inv); break;
                case 3: fft_bfly<3>(af, bf, tw, n, Ns, j, m, step, inv); break;
                case 4: fft_bfly<4>(af, bf, tw, n, Ns, j, m, step, inv); break;
                case 5: fft_bfly<5>(af, bf, tw, n, Ns, j, m, step, inv); break;
                default: fft_bfly<7>(af, bf, tw, n, Ns, j, m, step, inv); break;     // (radices 11, 13: the host keeps the GEMM chain)
            }
        }
        __syncthreads();
        cplx* t_ = a; a = b; b = t_;
        Ns *= R;
    }
    return a;
}
// ---------------------------------------------------------------------------------------------
// Specialised modem chain for 24-point transforms (the default geometry: 24 subcarriers; FBMC with TimeSpacing 12 and
// overlapping factor 8, CP-OFDM): the generic kernels above spend most of their issue slots on index arithmetic (the chain is
// issue-bound, not shared-memory-bound), so here every loop has compile-time trip counts and register-resident coefficients.
//   * 24-point DFT as 6 x 4 (Cooley-Tukey, b = 4 b1 + b2, m = m1 + 6 m2): pass 1 = one 6-point DFT over b1 per (symbol, b2) and
//     the twiddle e^{-+2 pi i b2 m1 / 24}; pass 2 = one 4-point DFT over b2 per (symbol, m1), in place.
//   * overlap-add (FBMC.m:267-268) as a 16-tap FIR along the symbol index: s[12 a + b] = 1/24 sum_j p[12 j + b] Z_{a-j}[b + 12 (j & 1)];
//     a work item owns one b and CH consecutive a: the 16 filter taps and the CH accumulators stay in registers, every
//     Z value is read once per item.
//   * fold (FBMC.m:297-302): a_k[m] = sum_o p[24 o + m] r[12 (k + 2 o + h) + b], m = b + 12 h; item = one m and CK consecutive k.
template <bool INV>
__device__ __forceinline__ void dft3_regs(cplx& x0, cplx& x1, cplx& x2) {
    const double h3 = 0.86602540378443864676;
    const cplx sm = cadd(x1, x2), df = csub(x1, x2);
    const cplx base = cmake(x0.x - 0.5 * sm.x, x0.y - 0.5 * sm.y);
    const cplx rot = INV ? cmake(-h3 * df.y, h3 * df.x) : cmake(h3 * df.y, -h3 * df.x);     // forward: -i h3 df
    x0 = cadd(x0, sm); x1 = cadd(base, rot); x2 = csub(base, rot);
}
// X[m] = sum_b x[b] e^{-+2 pi i b m / 6}: two 3-point DFTs over the even / odd inputs, then X[m] = E[m mod 3] + w6^m O[m mod 3]
template <bool INV>
__device__ __forceinline__ void dft6_regs(cplx (&x)[6]) {
    cplx e0 = x[0], e1 = x[2], e2 = x[4], o0 = x[1], o1 = x[3], o2 = x[5];
    dft3_regs<INV>(e0, e1, e2);
    dft3_regs<INV>(o0, o1, o2);
    const double c = 0.5, sn = INV ? 0.86602540378443864676 : -0.86602540378443864676;     // w6 = c + i sn
    const cplx t1 = cmake(c * o1.x - sn * o1.y, c * o1.y + sn * o1.x);                      // w6   o1
    const cplx t2 = cmake(-c * o2.x - sn * o2.y, -c * o2.y + sn * o2.x);                    // w6^2 o2 = (-c + i sn) o2
    x[0] = cadd(e0, o0); x[3] = csub(e0, o0);                  // w6^3 = -1
    x[1] = cadd(e1, t1); x[4] = csub(e1, t1);                  // w6^4 = -w6
    x[2] = cadd(e2, t2); x[5] = csub(e2, t2);                  // w6^5 = -w6^2
}
template <bool INV>
__device__ __forceinline__ void dft4_regs(cplx (&x)[4]) {
    const cplx s02 = cadd(x[0], x[2]), d02 = csub(x[0], x[2]), s13 = cadd(x[1], x[3]), d13 = csub(x[1], x[3]);
    const cplx r = INV ? cmake(-d13.y, d13.x) : cmake(d13.y, -d13.x);
    x[0] = cadd(s02, s13); x[1] = cadd(d02, r); x[2] = csub(s02, s13); x[3] = csub(d02, r);
}
// n_sym transforms of 24 points: A (input, destroyed) -> B (result), tw[m] = e^{-2 pi i m / 24}.  Ends with a block barrier.
template <bool INV>
__device__ __forceinline__ void fft24_batch(const cplx* __restrict__ A, cplx* __restrict__ B, const cplx* __restrict__ tw, int n_sym) {
    for (int item = threadIdx.x; item < n_sym * 4; item += blockDim.x) {
        const int k = item >> 2, b2 = item & 3;
        const cplx* a = A + k * 24 + b2;
        cplx x[6];
#pragma unroll
        for (int b1 = 0; b1 < 6; ++b1) x[b1] = a[4 * b1];
        dft6_regs<INV>(x);
        cplx* o = B + k * 24 + 6 * b2;
        o[0] = x[0];
#pragma unroll
        for (int m1 = 1; m1 < 6; ++m1) {
            cplx w = tw[b2 * m1];
            if (INV) w.y = -w.y;
            o[m1] = cmul(x[m1], w);
        }
    }
    __syncthreads();
    for (int item = threadIdx.x; item < n_sym * 6; item += blockDim.x) {
        const int k = item / 6, m1 = item - k * 6;
        cplx* a = B + k * 24 + m1;
        cplx x[4];
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2) x[b2] = a[6 * b2];
        dft4_regs<INV>(x);
#pragma unroll
        for (int m2 = 0; m2 < 4; ++m2) a[6 * m2] = x[m2];
    }
    __syncthreads();
}
// FBMC overlap-add for nfft = 24, TimeSpacing = 12, O = 8: Z [Ksym][24] -> S [N].  No trailing barrier.
template <int CH>
__device__ __forceinline__ void fbmc_overlap_add24(const cplx* __restrict__ Z, cplx* __restrict__ S, const double* __restrict__ filt,
                                                   int Ksym, int N) {
    const int A = (N + 11) / 12, nchunk = (A + CH - 1) / CH;
    for (int item = threadIdx.x; item < 12 * nchunk; item += blockDim.x) {
        const int ch = item / 12, b = item - ch * 12, a0 = ch * CH;
        double c[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) c[j] = filt[12 * j + b];
        cplx acc[CH];
#pragma unroll
        for (int ai = 0; ai < CH; ++ai) acc[ai] = cmake(0.0, 0.0);
#pragma unroll
        for (int d = 0; d < 15 + CH; ++d) {
            const int kk = a0 - 15 + d;
            if (kk < 0 || kk >= Ksym) continue;
            const cplx zA = Z[kk * 24 + b], zB = Z[kk * 24 + b + 12];
#pragma unroll
            for (int ai = 0; ai < CH; ++ai) {
                const int j = 15 + ai - d;                      // compile-time after unrolling
                if (j < 0 || j > 15) continue;
                const cplx z = (j & 1) ? zB : zA;
                acc[ai].x = fma(c[j], z.x, acc[ai].x); acc[ai].y = fma(c[j], z.y, acc[ai].y);
            }
        }
#pragma unroll
        for (int ai = 0; ai < CH; ++ai) {
            const int nn = 12 * (a0 + ai) + b;
            if (nn < N) S[nn] = cmake(acc[ai].x * (1.0 / 24), acc[ai].y * (1.0 / 24));
        }
    }
}
// FBMC fold for nfft = 24, TimeSpacing = 12, O = 8: R [N] -> Aout [Ksym][24].  No trailing barrier.
template <int CK>
__device__ __forceinline__ void fbmc_fold24(const cplx* __restrict__ R, cplx* __restrict__ Aout, const double* __restrict__ filt,
                                            int Ksym, int N) {
    const int nchunk = (Ksym + CK - 1) / CK;
    for (int item = threadIdx.x; item < 24 * nchunk; item += blockDim.x) {
        const int ch = item / 24, m = item - ch * 24, k0 = ch * CK, h = m >= 12 ? 1 : 0, b = m - 12 * h;
        double c[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) c[o] = filt[24 * o + m];
        cplx acc[CK];
#pragma unroll
        for (int ki = 0; ki < CK; ++ki) acc[ki] = cmake(0.0, 0.0);
#pragma unroll
        for (int d = 0; d < CK + 14; ++d) {
            const int idx = 12 * (k0 + h + d) + b;
            if (idx >= N) continue;
            const cplx r = R[idx];
#pragma unroll
            for (int ki = 0; ki < CK; ++ki) {
                const int e = d - ki;                           // = 2 o
                if (e < 0 || (e & 1) || e > 14) continue;
                acc[ki].x = fma(c[e / 2], r.x, acc[ki].x); acc[ki].y = fma(c[e / 2], r.y, acc[ki].y);
            }
        }
#pragma unroll
        for (int ki = 0; ki < CK; ++ki) if (k0 + ki < Ksym) Aout[(k0 + ki) * 24 + m] = acc[ki];
    }
}
// The chain  symbols in X0 [Ksym][24] (already phase-shifted and scaled)  ->  demodulated bins Y [Ksym][24]  with the banded
// channel `taps` ([T][N], tap t delayed by tap_delay[t]) in between.  Returns Y (= X1); X0 is free afterwards.
// Requires md.nfft == 24 and, for FBMC, time_spacing == 12, O == 8.  The caller has synchronised after filling X0.
template <int NH>
__device__ __forceinline__ cplx* modem_chain24(const ModemDev& md, cplx* X0, cplx* X1, const cplx* tw, const double* filt,
                                               const cplx* __restrict__ taps, int T, const int* __restrict__ tap_delay, int N) {
    const int Ksym = md.Ksym, TS = md.time_spacing, tid = threadIdx.x, nthr = blockDim.x;
    fft24_batch<true>(X0, X1, tw, Ksym);                                   // Z in X1
    if (md.kind == 0) fbmc_overlap_add24<5>(X1, X0, filt, Ksym, N);        // s in X0
    else
        for (int nn = tid; nn < N; nn += nthr) {                           // cyclic prefix + zero guards (OFDM.m:158-164)
            const int q = nn - md.zero_guard, k = q >= 0 ? q / TS : Ksym;
            cplx acc = cmake(0.0, 0.0);
            if (k < Ksym) { int m = q - k * TS - md.cp; if (m < 0) m += 24; acc = X1[k * 24 + m]; }
            X0[nn] = cmake(acc.x * (1.0 / 24), acc.y * (1.0 / 24));
        }
    __syncthreads();
    if (NH > 0) {                                                          // r = H s, into X1: N <= NH * blockDim.x; the tap loads of a
        cplx acc[NH > 0 ? NH : 1];                                         // thread's samples are issued together
#pragma unroll
        for (int u = 0; u < NH; ++u) acc[u] = cmake(0.0, 0.0);
        for (int t = 0; t < T; ++t) {
            const int d = tap_delay[t];
            const cplx* tp = taps + (int64_t)t * N;
            cplx hv[NH > 0 ? NH : 1];
#pragma unroll
            for (int u = 0; u < NH; ++u) { const int nn = tid + u * nthr; hv[u] = nn < N ? ld_nc(tp + nn) : cmake(0.0, 0.0); }
#pragma unroll
            for (int u = 0; u < NH; ++u) { const int nn = tid + u * nthr; if (nn < N && nn >= d) cfma(acc[u], hv[u], X0[nn - d]); }
        }
#pragma unroll
        for (int u = 0; u < NH; ++u) { const int nn = tid + u * nthr; if (nn < N) X1[nn] = acc[u]; }
    } else
    for (int nn = tid; nn < N; nn += nthr) {
        cplx acc = cmake(0.0, 0.0);
        for (int t = 0; t < T; ++t) {
            const int d = tap_delay[t];
            if (nn >= d) cfma(acc, ld_nc(taps + (int64_t)t * N + nn), X0[nn - d]);
        }
        X1[nn] = acc;
    }
    __syncthreads();
    if (md.kind == 0) fbmc_fold24<6>(X1, X0, filt, Ksym, N);               // folded symbols in X0
    else
        for (int idx = tid; idx < Ksym * 24; idx += nthr) {                // drop the cyclic prefix (OFDM.m:172-176)
            const int k = idx / 24, m = idx - k * 24;
            X0[idx] = X1[md.zero_guard + k * TS + md.cp + m];
        }
    __syncthreads();
    fft24_batch<false>(X0, X1, tw, Ksym);                                  // Y in X1
    return X1;
}
__host__ __device__ __forceinline__ bool modem_fast24(const ModemDev& md) {
    return md.nfft == 24 && (md.kind == 1 || (md.time_spacing == 12 && md.O == 8 && md.Np == 192));
}
// Head and tail of a chain column with every global operand of a thread's NE symbols loaded in one batch (the chain kernels
// are latency-bound: independent loads in flight are what hides it).  K <= NE * blockDim.x.
//   head: X0[symbol][bin] = v * phase * norm (the modulator's input),  e = y + h v  (kept in registers across the chain)
//   tail: out = e - conj(phase) Y[bin] / (norm F)                      (divided by hdiv where the caller equalises right away)
#define CHAIN24_NE 6
#ifndef CHAIN24_EST_NB
#define CHAIN24_EST_NB 3
#endif
#ifndef CHAIN24_EST_MIN_CTAS
#define CHAIN24_EST_MIN_CTAS CHAIN24_MIN_CTAS
#endif
#ifndef CHAIN24_MIN_CTAS
#define CHAIN24_MIN_CTAS 6          // 80 registers; measured against 5 (102 registers) and 4 (128): see profiles
#endif
#define CHAIN24_NH 5
// (register-carried variant: e = y + h v stays in NE complex registers across the chain -- measured faster for k_est_factored,
// whose three strided operand streams want all 4 NE loads in flight; the shared-memory variant below is faster for k_perfect_fbmc_det)
template <int NE>
__device__ __forceinline__ void chain24_head_regs(const ModemDev& md, cplx* X0, const cplx* __restrict__ vcol, int vstride, const cplx* __restrict__ hcol,
                                             int hstride, const cplx* __restrict__ ycol, int K, cplx (&e)[NE]) {
    const int tid = threadIdx.x, nthr = blockDim.x, L = md.L;
    const bool fbmc = md.kind == 0;
    cplx vv[NE], hh[NE], ph[NE];
    int bl[NE];
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        const int i = tid + u * nthr;
        const bool ok = i < K;
        vv[u] = ok ? ld_nc(vcol + (int64_t)i * vstride) : cmake(0.0, 0.0);
        hh[u] = ok ? ld_nc(hcol + (int64_t)i * hstride) : cmake(0.0, 0.0);
        e[u] = ok ? ld_nc(ycol + i) : cmake(0.0, 0.0);
        ph[u] = (ok && fbmc) ? ld_nc(md.phase + i) : cmake(1.0, 0.0);
        bl[u] = ok ? md.bin[i % L] : 0;
    }
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        const int i = tid + u * nthr;
        if (i >= K) continue;
        cfma(e[u], hh[u], vv[u]);
        const cplx x = cmul(vv[u], ph[u]);
        X0[(i / L) * 24 + bl[u]] = cmake(x.x * md.norm, x.y * md.norm);
    }
}
template <int NE, bool DIV>
__device__ __forceinline__ void chain24_tail_regs(const ModemDev& md, const cplx* Y, const int* bins, const cplx (&e)[NE],
                                             const cplx* __restrict__ hdiv, cplx* out, int ostride, int K) {
    const int tid = threadIdx.x, nthr = blockDim.x, L = md.L;
    const bool fbmc = md.kind == 0;
    cplx ph[NE], hd[NE];
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        const int i = tid + u * nthr;
        const bool ok = i < K;
        ph[u] = (ok && fbmc) ? ld_nc(md.phase + i) : cmake(1.0, 0.0);
        if (DIV) hd[u] = ok ? ld_nc(hdiv + i) : cmake(1.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        const int i = tid + u * nthr;
        if (i >= K) continue;
        const int k = i / L, l = i - k * L;
        const cplx u0 = cmulc(ph[u], Y[k * 24 + bins[l]]);
        const cplx r = cmake(e[u].x - u0.x * md.inv_demod, e[u].y - u0.y * md.inv_demod);
        out[(int64_t)i * ostride] = DIV ? cdiv_fast(r, hd[u]) : r;
    }
}

template <int NE, int NB>
__device__ __forceinline__ void chain24_head(const ModemDev& md, cplx* X0, const cplx* __restrict__ vcol, int vstride, const cplx* __restrict__ hcol,
                                             int hstride, const cplx* __restrict__ ycol, int K, cplx* __restrict__ E) {
    const int tid = threadIdx.x, nthr = blockDim.x, L = md.L;
    const bool fbmc = md.kind == 0;
    // NB symbols per batch: 4 x NB 16-byte loads in flight per thread
#pragma unroll
    for (int u0 = 0; u0 < NE; u0 += NB) {
        cplx vv[NB], hh[NB], ph[NB], e[NB];
        int bl[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int i = tid + (u0 + u) * nthr;
            const bool ok = i < K;
            vv[u] = ok ? ld_nc(vcol + (int64_t)i * vstride) : cmake(0.0, 0.0);
            hh[u] = ok ? ld_nc(hcol + (int64_t)i * hstride) : cmake(0.0, 0.0);
            e[u] = ok ? ld_nc(ycol + i) : cmake(0.0, 0.0);
            ph[u] = (ok && fbmc) ? ld_nc(md.phase + i) : cmake(1.0, 0.0);
            bl[u] = ok ? md.bin[i % L] : 0;
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int i = tid + (u0 + u) * nthr;
            if (i >= K) continue;
            cfma(e[u], hh[u], vv[u]);
            E[i] = e[u];                                        // (carried in registers it spilled: shared memory instead)
            const cplx x = cmul(vv[u], ph[u]);
            X0[(i / L) * 24 + bl[u]] = cmake(x.x * md.norm, x.y * md.norm);
        }
    }
}
template <int NE, bool DIV>
__device__ __forceinline__ void chain24_tail(const ModemDev& md, const cplx* Y, const int* bins, const cplx* __restrict__ E,
                                             const cplx* __restrict__ hdiv, cplx* out, int ostride, int K) {
    const int tid = threadIdx.x, nthr = blockDim.x, L = md.L;
    const bool fbmc = md.kind == 0;
    cplx ph[NE], hd[NE];
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        const int i = tid + u * nthr;
        const bool ok = i < K;
        ph[u] = (ok && fbmc) ? ld_nc(md.phase + i) : cmake(1.0, 0.0);
        if (DIV) hd[u] = ok ? ld_nc(hdiv + i) : cmake(1.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        const int i = tid + u * nthr;
        if (i >= K) continue;
        const int k = i / L, l = i - k * L;
        const cplx u0 = cmulc(ph[u], Y[k * 24 + bins[l]]);
        const cplx ev = E[i];
        const cplx r = cmake(ev.x - u0.x * md.inv_demod, ev.y - u0.y * md.inv_demod);
        out[(int64_t)i * ostride] = DIV ? cdiv_fast(r, hd[u]) : r;
    }
}

template <int CW>
__global__ void __launch_bounds__(PERF_FBMC_THREADS, PERF_FBMC_MIN_CTAS) k_perfect_fbmc(PerfFbmcParams p) {
    extern __shared__ __align__(16) cplx pf_smem[];
    const ModemDev& md = p.md;
    const int n = md.nfft, Ksym = md.Ksym, L = md.L, N = p.N, K = p.K, TS = md.time_spacing, nx = Ksym * n;
    // two buffers of CW x Ksym x n values (>= CW x N) ping-pong through the chain:
    //   X0: padded symbols -> (IFFT) Z in Xz -> (overlap-add) s in Xo -> (H) r in Xz -> (fold) in Xo -> (FFT) Y
    cplx* X0 = pf_smem;
    cplx* X1 = X0 + CW * nx;
    cplx* tw = X1 + CW * nx;               // [n]
    double* filt = reinterpret_cast<double*>(tw + n);          // [Np] prototype filter
    int* bins = reinterpret_cast<int*>(filt + md.Np);          // [L]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int2 grp = p.groups[blockIdx.x];
    const int c0 = grp.x, nc = grp.y;
    for (int m = tid; m < n; m += nthr) tw[m] = md.tw[m];
    for (int m = tid; m < md.Np; m += nthr) filt[m] = md.filt[m];
    for (int m = tid; m < L; m += nthr) bins[m] = md.bin[m];
    if (L < n) for (int idx = tid; idx < CW * nx; idx += nthr) X0[idx] = cmake(0.0, 0.0);      // bins without a subcarrier
    __syncthreads();
    // ---- load v: the CW columns are neighbours in the unit scratch (consecutive threads, consecutive 16-byte slots)
    for (int idx = tid; idx < CW * K; idx += nthr) {            // flat index: L may be far below the thread count
        const int c = idx % CW, i = idx / CW, k = i / L, l = i - k * L;
        cplx v = cmake(0.0, 0.0);
        if (c < nc) {
            v = cmul(p.v_base[p.voff[c0 + c] + (int64_t)i * NC_MAX], md.phase[i]);
            v = cmake(v.x * md.norm, v.y * md.norm);
        }
        X0[(c * Ksym + k) * n + bins[l]] = v;
    }
    __syncthreads();
    cplx* Xz = fft_shared_batch(X0, X1, tw, md.plan, true, CW * Ksym);              // un-normalised IFFT per symbol
    cplx* Xo = (Xz == X0) ? X1 : X0;
    // ---- overlap-add: s[nn] = 1/n sum_k p[nn - k TS] Z_k[(nn - k TS) mod n]
    const double inv_n = 1.0 / n;
    for (int c = 0; c < CW; ++c)
        for (int nn = tid; nn < N; nn += nthr) {
            int k_lo = (nn - md.Np + TS) / TS; if (nn - md.Np + 1 <= 0) k_lo = 0;
            const int k_hi = min(Ksym - 1, nn / TS);
            cplx acc = cmake(0.0, 0.0);
            int tap = nn - k_lo * TS, mm = tap % n;
            const cplx* zc = Xz + (c * Ksym + k_lo) * n;
            for (int k = k_lo; k <= k_hi; ++k) {
                if (tap >= 0 && tap < md.Np) {
                    const cplx z = zc[mm];
                    const double pf = filt[tap];
                    acc.x = fma(pf, z.x, acc.x); acc.y = fma(pf, z.y, acc.y);
                }
                tap -= TS; mm -= TS; if (mm < 0) mm += n;
                zc += n;
            }
            Xo[c * N + nn] = cmake(acc.x * inv_n, acc.y * inv_n);
        }
    __syncthreads();
    // ---- r = H s with the realization's taps (Z is dead: r goes into its buffer)
    for (int c = 0; c < CW; ++c) {
        const cplx* hr = p.h + (int64_t)p.rep[c0 + min(c, nc - 1)] * p.T * N;
        for (int nn = tid; nn < N; nn += nthr) {
            cplx acc = cmake(0.0, 0.0);
            for (int t = 0; t < p.T; ++t) {
                const int d = p.tap_delay[t];
                if (nn >= d) cfma(acc, hr[(int64_t)t * N + nn], Xo[c * N + nn - d]);
            }
            Xz[c * N + nn] = acc;
        }
    }
    __syncthreads();
    // ---- fold: a_k[m] = sum_o p[o n + m] r[k TS + o n + m]   (s is dead: into its buffer)
    for (int idx = tid; idx < CW * nx; idx += nthr) {
        const int c = idx / nx, rem = idx - c * nx, k = rem / n, m = rem - k * n;
        const cplx* seg = Xz + c * N + k * TS + m;
        cplx acc = cmake(0.0, 0.0);
        for (int o = 0; o < md.O; ++o) {
            const double pf = filt[o * n + m];
            const cplx v = seg[o * n];
            acc.x = fma(pf, v.x, acc.x); acc.y = fma(pf, v.y, acc.y);
        }
        Xo[idx] = acc;
    }
    __syncthreads();
    const cplx* Y = fft_shared_batch(Xo, Xz, tw, md.plan, false, CW * Ksym);
    // ---- epilogue: y_ic = y - conj(phase) Y[bin] / (norm F) + h v
    for (int idx = tid; idx < CW * K; idx += nthr) {
        const int c = idx % CW, i = idx / CW;
        if (c >= nc) continue;
        const int k = i / L, l = i - k * L, col = c0 + c;
        const cplx u0 = cmulc(md.phase[i], Y[(c * Ksym + k) * n + bins[l]]);
        const int64_t o = p.voff[col] + (int64_t)i * NC_MAX;
        const cplx yv = p.y[p.yoff[col] + i], hv = p.htrue[(int64_t)p.rep[col] * K + i], vv = p.v_base[o];
        const cplx hvv = cmul(hv, vv);
        p.y_base[o] = cmake(yv.x - u0.x * md.inv_demod + hvv.x, yv.y - u0.y * md.inv_demod + hvv.y);
    }
}

// k_perfect_fbmc for one column per CTA with the rest of the perfect-CSI iteration behind it: instead of writing y_ic the
// epilogue equalises (x = y_ic / h, DS.m:545-548), a block barrier later the data symbols are selected / de-spread, decided
// and counted exactly as k_ic_light does (DS.m:549-561), and the decided words go to a byte array from which k_ic_light
// builds the next v = C z.  it = 0 is the one-tap stage (x = y / h, DS.m:450-466): no modem chain.
struct PerfDetParams {
    ModemDev md;
    SchemeDev sch[2]; ConstDev cst[2]; int scheme_id[2];
    int it, n_iter, n_snr, n_rep, nsch, n_cols, T, N, K, zw_stride;
    const int64_t* voff; const int64_t* yoff; const int* rep;
    const cplx* v_base; const cplx* y; const cplx* htrue; const cplx* h; const int* tap_delay;
    uint8_t* zw_g; uint32_t* err;
    int vstride;                   // distance of a column's consecutive rows in the unit scratch: NC_MAX (row-major) or 1 (column-major)
};
template <bool FAST24>
__global__ void __launch_bounds__(PERF_FBMC_THREADS, FAST24 ? CHAIN24_MIN_CTAS : PERF_FBMC_MIN_CTAS) k_perfect_fbmc_det(PerfDetParams p) {
    extern __shared__ __align__(16) cplx pf_smem[];
    const ModemDev& md = p.md;
    const int n = md.nfft, Ksym = md.Ksym, L = md.L, N = p.N, K = p.K, TS = md.time_spacing, nx = Ksym * n;
    const int nbuf = max(nx, N);                               // CP-OFDM with zero guards: N can exceed Ksym * nfft
    const bool fbmc = md.kind == 0;
    cplx* X0 = pf_smem;
    cplx* X1 = X0 + nbuf;
    cplx* tw = X1 + nbuf;
    double* filt = reinterpret_cast<double*>(tw + n);
    int* bins = reinterpret_cast<int*>(filt + md.Np);
    __shared__ unsigned int cnt[2];
    const int tid = threadIdx.x, nthr = blockDim.x, col = blockIdx.x;
    const int snr = col % p.n_snr, slot = (col / p.n_snr) % p.nsch, rep = p.rep[col];
    const SchemeDev& sd = p.sch[slot];
    const ConstDev& cd = p.cst[sd.constellation];
    const cplx* ycol = p.y + p.yoff[col];
    const cplx* ht = p.htrue + (int64_t)rep * K;
    if (tid < 2) cnt[tid] = 0;
    cplx* Xe;
    if (p.it == 0) {
        for (int i = tid; i < K; i += nthr) X1[i] = cdiv_fast(ycol[i], ht[i]);
        Xe = X1;
    } else if (FAST24) {                                       // K <= CHAIN24_NE * 128, N <= CHAIN24_NH * 128 (the host checks)
        if (tid < 24) tw[tid] = md.tw[tid];
        if (fbmc) for (int m = tid; m < md.Np; m += nthr) filt[m] = md.filt[m];
        if (tid < L) bins[tid] = md.bin[tid];
        if (L < 24) { for (int idx = tid; idx < nx; idx += nthr) X0[idx] = cmake(0.0, 0.0); __syncthreads(); }
        cplx* e = reinterpret_cast<cplx*>((reinterpret_cast<uintptr_t>(bins + L) + 15) & ~(uintptr_t)15);
        chain24_head<CHAIN24_NE, 3>(md, X0, p.v_base + p.voff[col], p.vstride, ht, 1, ycol, K, e);
        __syncthreads();
        const cplx* Y = modem_chain24<CHAIN24_NH>(md, X0, X1, tw, filt, p.h + (int64_t)rep * p.T * N, p.T, p.tap_delay, N);
        chain24_tail<CHAIN24_NE, true>(md, Y, bins, e, ht, X0, 1, K);
        Xe = X0;
    } else {
        for (int m = tid; m < n; m += nthr) tw[m] = md.tw[m];
        if (fbmc) for (int m = tid; m < md.Np; m += nthr) filt[m] = md.filt[m];
        for (int m = tid; m < L; m += nthr) bins[m] = md.bin[m];
        if (L < n) for (int idx = tid; idx < nx; idx += nthr) X0[idx] = cmake(0.0, 0.0);
        __syncthreads();
        const cplx* vcol = p.v_base + p.voff[col];
        for (int i = tid; i < K; i += nthr) {
            const int k = i / L, l = i - k * L;
            cplx v = vcol[(int64_t)i * p.vstride];
            if (fbmc) v = cmul(v, md.phase[i]);
            X0[k * n + bins[l]] = cmake(v.x * md.norm, v.y * md.norm);
        }
        __syncthreads();
        cplx* Y;
        {
            cplx* Xz = fft_shared_batch(X0, X1, tw, md.plan, true, Ksym);
            cplx* Xo = (Xz == X0) ? X1 : X0;
            const double inv_n = 1.0 / n;
            for (int nn = tid; nn < N; nn += nthr) {
                cplx acc = cmake(0.0, 0.0);
                if (fbmc) {                                         // overlap-add (FBMC.m:267-268)
                    int k_lo = (nn - md.Np + TS) / TS; if (nn - md.Np + 1 <= 0) k_lo = 0;
                    const int k_hi = min(Ksym - 1, nn / TS);
                    int tap = nn - k_lo * TS, mm = tap % n;
                    const cplx* zc = Xz + k_lo * n;
                    for (int k = k_lo; k <= k_hi; ++k) {
                        if (tap >= 0 && tap < md.Np) {
                            const cplx z = zc[mm];
                            const double pf = filt[tap];
                            acc.x = fma(pf, z.x, acc.x); acc.y = fma(pf, z.y, acc.y);
                        }
                        tap -= TS; mm -= TS; if (mm < 0) mm += n;
                        zc += n;
                    }
                } else {                                            // cyclic prefix + zero guards (OFDM.m:158-164)
                    const int q = nn - md.zero_guard, k = q >= 0 ? q / TS : Ksym;
                    if (k < Ksym) { int m = q - k * TS - md.cp; if (m < 0) m += n; acc = Xz[k * n + m]; }
                }
                Xo[nn] = cmake(acc.x * inv_n, acc.y * inv_n);
            }
            __syncthreads();
            const cplx* hr = p.h + (int64_t)rep * p.T * N;
            for (int nn = tid; nn < N; nn += nthr) {                // r = H s
                cplx acc = cmake(0.0, 0.0);
                for (int t = 0; t < p.T; ++t) {
                    const int d = p.tap_delay[t];
                    if (nn >= d) cfma(acc, hr[(int64_t)t * N + nn], Xo[nn - d]);
                }
                Xz[nn] = acc;
            }
            __syncthreads();
            for (int idx = tid; idx < nx; idx += nthr) {            // fold by O (FBMC.m:297-302) / drop the cyclic prefix (OFDM.m:172-176)
                const int k = idx / n, m = idx - k * n;
                cplx acc = cmake(0.0, 0.0);
                if (fbmc) {
                    const cplx* seg = Xz + k * TS + m;
                    for (int o = 0; o < md.O; ++o) {
                        const double pf = filt[o * n + m];
                        const cplx v = seg[o * n];
                        acc.x = fma(pf, v.x, acc.x); acc.y = fma(pf, v.y, acc.y);
                    }
                } else acc = Xz[md.zero_guard + k * TS + md.cp + m];
                Xo[idx] = acc;
            }
            __syncthreads();
            Y = fft_shared_batch(Xo, Xz, tw, md.plan, false, Ksym);
            Xe = (Y == X0) ? X1 : X0;
        }
        for (int i = tid; i < K; i += nthr) {                   // y_ic = y - U + h v, x = y_ic / h
            const int k = i / L, l = i - k * L;
            cplx u0 = Y[k * n + bins[l]];
            if (fbmc) u0 = cmulc(md.phase[i], u0);
            const cplx hv = ht[i], yv = ycol[i];
            const cplx hvv = cmul(hv, vcol[(int64_t)i * p.vstride]);
            Xe[i] = cdiv_fast(cmake(yv.x - u0.x * md.inv_demod + hvv.x, yv.y - u0.y * md.inv_demod + hvv.y), hv);
        }
    }
    __syncthreads();
    const int P = sd.P, n_data = sd.n_data;
    const bool select = sd.detect_mode != 1, last = p.it == p.n_iter;
    const uint32_t* txw = sd.txword + (int64_t)rep * n_data;
    const int64_t colbase = ((int64_t)snr * p.n_rep + rep) * n_data;
    uint8_t* zw = p.zw_g + (int64_t)col * p.zw_stride;
    unsigned e_all = 0, e_edge = 0;
    for (int d = tid; d < n_data; d += nthr) {
        cplx xd;
        if (select) {
            const cplx xh = Xe[sd.data_pos[d]];
            xd = cmake(xh.x * sd.inv_sqrt_dpr, sd.detect_mode == 0 ? 0.0 : xh.y * sd.inv_sqrt_dpr);
        } else {
            cplx acc = cmake(0.0, 0.0);
            for (int e = sd.ct_colptr[P + d]; e < sd.ct_colptr[P + d + 1]; ++e) {
                const cplx t = cmulc(sd.ct_val[e], Xe[sd.ct_row[e]]);
                acc.x += t.x; acc.y += t.y;
            }
            xd = cmake(acc.x * sd.inv_dpr, 0.0);
        }
        const int word = ic_decide(cd, xd, txw[d], sd.edge_mask[d], e_all, e_edge);
        if (last) sd.xD[1][colbase + d] = xd;
        else zw[d] = (uint8_t)word;
    }
    if (e_all) atomicAdd(&cnt[0], e_all);
    if (e_edge) atomicAdd(&cnt[1], e_edge);
    __syncthreads();
    if (tid < 2)
        p.err[((((int64_t)rep * p.n_snr + snr) * (p.n_iter + 1) + p.it) * 3 + p.scheme_id[slot]) * 4 + 2 + tid] = cnt[tid];
}

// ---------------------------------------------------------------------------------------------
// The whole perfect-CSI twin of an FBMC column (DS.m:450-466 and 541-561 for every iteration) in ONE launch per batch:
// one CTA owns one (realization, scheme, SNR point) column and runs, entirely in shared memory,
//     it = 0:        x = y / h                                   -> select / de-spread, decide, count
//     it = 1..I:     v = C [xP; decided symbols]                 (precoder in CSR form)
//                    y_ic = y - Demodulation(H Modulation(v)) + h v     (polyphase modem, as k_perfect_fbmc)
//                    x = y_ic / h                                -> select / de-spread, decide, count
// The twin needs nothing from the estimated-CSI path, so v and y_ic never touch HBM and the scalar kernel k_ic_light has no
// perfect-CSI units left for this waveform.  Decisions, counters and the parity state (chest_get_state) follow k_ic_light.
// Opt-in (CHEST_TWIN=1): on B200 it measured SLOWER than PERF units + k_perfect_fbmc (20.9 vs 15.8 ms per step at the default
// geometry) because every column re-reads the long precoder rows that k_ic_light shares across 16 columns on DMMA tiles.
struct PerfTwinParams {
    ModemDev md;
    SchemeDev sch[2];              // the schemes on this waveform, by slot
    ConstDev cst[2];
    int scheme_id[2];              // their global ids (index of the error counters)
    int nsch, n_snr, n_rep, n_iter, T, N, K, n_long_max;
    const cplx* htrue; const cplx* h; const int* tap_delay;
    uint32_t* err;
};
__global__ void __launch_bounds__(PERF_FBMC_THREADS, PERF_FBMC_MIN_CTAS) k_perfect_twin_fbmc(PerfTwinParams p) {
    extern __shared__ __align__(16) cplx pf_smem[];
    const ModemDev& md = p.md;
    const int n = md.nfft, Ksym = md.Ksym, L = md.L, N = p.N, K = p.K, TS = md.time_spacing, nx = Ksym * n;
    cplx* X0 = pf_smem;
    cplx* X1 = X0 + nx;
    cplx* tw = X1 + nx;
    double* filt = reinterpret_cast<double*>(tw + n);
    int* bins = reinterpret_cast<int*>(filt + md.Np);
    cplx* vlong = reinterpret_cast<cplx*>(bins + ((L + 3) & ~3));   // v of the long precoder rows (auxiliary / spread positions)
    unsigned short* lrmap = reinterpret_cast<unsigned short*>(vlong + p.n_long_max);   // row -> index into vlong
    uint8_t* zw = reinterpret_cast<uint8_t*>(lrmap + ((K + 7) & ~7));                  // decided word per data symbol
    __shared__ unsigned int cnt[2];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int col = blockIdx.x, snr = col % p.n_snr, slot = (col / p.n_snr) % p.nsch, rep = col / (p.n_snr * p.nsch);
    const SchemeDev& sd = p.sch[slot];
    const ConstDev& cd = p.cst[sd.constellation];
    const int P = sd.P, n_data = sd.n_data;
    const bool select = sd.detect_mode != 1;
    const int scheme_id = p.scheme_id[slot];
    const cplx* ycol = sd.y + ((int64_t)snr * p.n_rep + rep) * K;
    const cplx* ht = p.htrue + (int64_t)rep * K;
    const cplx* xP = sd.xP + (int64_t)rep * P;
    const uint32_t* txw = sd.txword + (int64_t)rep * n_data;
    const cplx* hr = p.h + (int64_t)rep * p.T * N;
    const int64_t colbase = ((int64_t)snr * p.n_rep + rep) * n_data;
    for (int m = tid; m < n; m += nthr) tw[m] = md.tw[m];
    for (int m = tid; m < md.Np; m += nthr) filt[m] = md.filt[m];
    for (int m = tid; m < L; m += nthr) bins[m] = md.bin[m];
    for (int r = tid; r < sd.n_long_rows; r += nthr) lrmap[sd.long_rows[r]] = (unsigned short)r;
    // z[c] of the precoder input: transmitted pilot, or the symbol decided for data index c - P
    auto z_of = [&](int c) -> cplx { return c < P ? xP[c] : cd.symbol[zw[c - P]]; };
    auto v_of = [&](int i) -> cplx {                            // v[i] = sum_c C[i, c] z[c]
        const int c0 = sd.row_col0[i];
        if (c0 >= 0) return cmul(sd.row_val0[i], z_of(c0));
        return c0 == -2 ? vlong[lrmap[i]] : cmake(0.0, 0.0);    // long rows: summed once per iteration, one warp per row
    };
    for (int it = 0; it <= p.n_iter; ++it) {
        const bool last = it == p.n_iter;
        if (tid < 2) cnt[tid] = 0;
        cplx* Xe;                                               // equalised symbols x = y_ic / h of this iteration, [K]
        if (it == 0) {
            __syncthreads();
            for (int i = tid; i < K; i += nthr) X1[i] = cdiv_fast(ycol[i], ht[i]);
            Xe = X1;
        } else {
            if (L < n) for (int idx = tid; idx < nx; idx += nthr) X0[idx] = cmake(0.0, 0.0);
            __syncthreads();                                    // decided words of the previous iteration are complete
            for (int r = tid >> 5; r < sd.n_long_rows; r += nthr >> 5) {
                const int i = sd.long_rows[r];
                cplx acc = cmake(0.0, 0.0);
                for (int e = sd.c_rowptr[i] + (tid & 31); e < sd.c_rowptr[i + 1]; e += 32) cfma(acc, sd.c_val[e], z_of(sd.c_col[e]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); }
                if ((tid & 31) == 0) vlong[r] = acc;
            }
            __syncthreads();
            for (int i = tid; i < K; i += nthr) {
                const int k = i / L, l = i - k * L;
                const cplx v = cmul(v_of(i), md.phase[i]);
                X0[k * n + bins[l]] = cmake(v.x * md.norm, v.y * md.norm);
            }
            __syncthreads();
            cplx* Xz = fft_shared_batch(X0, X1, tw, md.plan, true, Ksym);
            cplx* Xo = (Xz == X0) ? X1 : X0;
            const double inv_n = 1.0 / n;
            for (int nn = tid; nn < N; nn += nthr) {            // overlap-add
                int k_lo = (nn - md.Np + TS) / TS; if (nn - md.Np + 1 <= 0) k_lo = 0;
                const int k_hi = min(Ksym - 1, nn / TS);
                cplx acc = cmake(0.0, 0.0);
                int tap = nn - k_lo * TS, mm = tap % n;
                const cplx* zc = Xz + k_lo * n;
                for (int k = k_lo; k <= k_hi; ++k) {
                    if (tap >= 0 && tap < md.Np) {
                        const cplx z = zc[mm];
                        const double pf = filt[tap];
                        acc.x = fma(pf, z.x, acc.x); acc.y = fma(pf, z.y, acc.y);
                    }
                    tap -= TS; mm -= TS; if (mm < 0) mm += n;
                    zc += n;
                }
                Xo[nn] = cmake(acc.x * inv_n, acc.y * inv_n);
            }
            __syncthreads();
            for (int nn = tid; nn < N; nn += nthr) {            // r = H s
                cplx acc = cmake(0.0, 0.0);
                for (int t = 0; t < p.T; ++t) {
                    const int d = p.tap_delay[t];
                    if (nn >= d) cfma(acc, hr[(int64_t)t * N + nn], Xo[nn - d]);
                }
                Xz[nn] = acc;
            }
            __syncthreads();
            for (int idx = tid; idx < nx; idx += nthr) {        // fold
                const int k = idx / n, m = idx - k * n;
                const cplx* seg = Xz + k * TS + m;
                cplx acc = cmake(0.0, 0.0);
                for (int o = 0; o < md.O; ++o) {
                    const double pf = filt[o * n + m];
                    const cplx v = seg[o * n];
                    acc.x = fma(pf, v.x, acc.x); acc.y = fma(pf, v.y, acc.y);
                }
                Xo[idx] = acc;
            }
            __syncthreads();
            cplx* Y = fft_shared_batch(Xo, Xz, tw, md.plan, false, Ksym);
            Xe = (Y == X0) ? X1 : X0;                           // the FFT scratch buffer is free again
            for (int i = tid; i < K; i += nthr) {               // y_ic = y - U + h v, then x = y_ic / h
                const int k = i / L, l = i - k * L;
                const cplx u0 = cmulc(md.phase[i], Y[k * n + bins[l]]);
                const cplx hv = ht[i], yv = ycol[i];
                const cplx hvv = cmul(hv, v_of(i));
                const cplx yic = cmake(yv.x - u0.x * md.inv_demod + hvv.x, yv.y - u0.y * md.inv_demod + hvv.y);
                Xe[i] = cdiv_fast(yic, hv);
            }
        }
        __syncthreads();                                        // x complete; every v_of() of this iteration has read zw
        unsigned e_all = 0, e_edge = 0;
        for (int d = tid; d < n_data; d += nthr) {
            cplx xd;
            if (select) {
                const cplx xh = Xe[sd.data_pos[d]];
                xd = cmake(xh.x * sd.inv_sqrt_dpr, sd.detect_mode == 0 ? 0.0 : xh.y * sd.inv_sqrt_dpr);
            } else {                                            // de-spreading: x_d = C_d^H x / dpr over the spreading set
                cplx acc = cmake(0.0, 0.0);
                for (int e = sd.ct_colptr[P + d]; e < sd.ct_colptr[P + d + 1]; ++e) {
                    const cplx t = cmulc(sd.ct_val[e], Xe[sd.ct_row[e]]);
                    acc.x += t.x; acc.y += t.y;
                }
                xd = cmake(acc.x * sd.inv_dpr, 0.0);
            }
            const int word = ic_decide(cd, xd, txw[d], sd.edge_mask[d], e_all, e_edge);
            if (last) sd.xD[1][colbase + d] = xd;
            zw[d] = (uint8_t)word;
        }
        if (e_all) atomicAdd(&cnt[0], e_all);
        if (e_edge) atomicAdd(&cnt[1], e_edge);
        __syncthreads();
        if (tid < 2) {
            const int64_t o = ((((int64_t)rep * p.n_snr + snr) * (p.n_iter + 1) + it) * 3 + scheme_id) * 4 + 2 + tid;     // csi = 1 (perfect)
            p.err[o] = cnt[tid];
        }
    }
}

// y = Q^H (r0 + sqrt(Pn/2) noise) of FBMC columns through the polyphase demodulator (DS.m:401-409 with Q^H r = Demodulation(r),
// FBMC.m:287-302): column = (scheme slot g, SNR point, realization) as in k_gemm<GEMM_DEMOD>; one column per CTA, the received
// samples, the folded symbols and the FFT ping-pong in two shared-memory buffers.
struct DemodFbmcParams {
    ModemDev md;
    int N, K, n_snr, n_rep, n_cols, fast24;
    const cplx* r0; const cplx* noise; const double* noise_scale;
    cplx* y;                       // [col][K]
};
__global__ void __launch_bounds__(PERF_FBMC_THREADS, 6) k_demod_fbmc(DemodFbmcParams p) {
    extern __shared__ __align__(16) cplx pf_smem[];
    const ModemDev& md = p.md;
    const int n = md.nfft, Ksym = md.Ksym, L = md.L, N = p.N, K = p.K, TS = md.time_spacing, nx = Ksym * n;
    const int nbuf = max(nx, N);
    const bool fbmc = md.kind == 0;
    cplx* X0 = pf_smem;
    cplx* X1 = X0 + nbuf;
    cplx* tw = X1 + nbuf;
    double* filt = reinterpret_cast<double*>(tw + n);
    int* bins = reinterpret_cast<int*>(filt + md.Np);
    const int tid = threadIdx.x, nthr = blockDim.x, col = blockIdx.x;
    for (int m = tid; m < n; m += nthr) tw[m] = md.tw[m];
    if (fbmc) for (int m = tid; m < md.Np; m += nthr) filt[m] = md.filt[m];
    for (int m = tid; m < L; m += nthr) bins[m] = md.bin[m];
    {   // received samples of this column
        const int rep = col % p.n_rep, q = col / p.n_rep, snr = q % p.n_snr, grp = q / p.n_snr;
        const cplx* a = p.r0 + ((int64_t)grp * p.n_rep + rep) * N;
        const cplx* nz = p.noise + ((int64_t)rep * p.n_snr + snr) * N;
        const double sc = p.noise_scale[snr];
        for (int nn = tid; nn < N; nn += nthr) { const cplx x = a[nn], z = nz[nn]; X1[nn] = cmake(x.x + sc * z.x, x.y + sc * z.y); }
    }
    __syncthreads();
    const bool fast24 = modem_fast24(md) && p.fast24;          // 24-point transforms: the specialised fold / DFT passes (modem_chain24)
    if (fast24 && fbmc) fbmc_fold24<6>(X1, X0, filt, Ksym, N);
    else
    for (int idx = tid; idx < nx; idx += nthr) {               // fold: a_k[m] = sum_o p[o n + m] r[k TS + o n + m]; OFDM: drop the prefix
        const int k = idx / n, m = idx - k * n;
        cplx acc = cmake(0.0, 0.0);
        if (fbmc) {
            const cplx* seg = X1 + k * TS + m;
            for (int o = 0; o < md.O; ++o) {
                const double pf = filt[o * n + m];
                const cplx v = seg[o * n];
                acc.x = fma(pf, v.x, acc.x); acc.y = fma(pf, v.y, acc.y);
            }
        } else acc = X1[md.zero_guard + k * TS + md.cp + m];
        X0[idx] = acc;
    }
    __syncthreads();
    const cplx* Y;
    if (fast24) { fft24_batch<false>(X0, X1, tw, Ksym); Y = X1; }
    else Y = fft_shared_batch(X0, X1, tw, md.plan, false, Ksym);
    cplx* out = p.y + (int64_t)col * K;
    for (int i = tid; i < K; i += nthr) {
        const int k = i / L, l = i - k * L;
        cplx u0 = Y[k * n + bins[l]];
        if (fbmc) u0 = cmulc(md.phase[i], u0);
        out[i] = cmake(u0.x * md.inv_demod, u0.y * md.inv_demod);
    }
}

// ---------------------------------------------------------------------------------------------
// Factored estimated-CSI cancellation.  The MMSE estimate of the transmission matrix is linear in the pilot estimates,
//     D-hat = sum_p W_p hP(p),   W = R_Dij_hP pinv(R_hP_est),   R_Dij_hP(:, q) = vec(Q^H M_q G)        (DS.m:256-289, 417-425)
// with M_q the banded pseudo-channel of pilot q (DS.m:260), so  D-hat = Q^H H-hat G  with the ESTIMATED CHANNEL
//     H-hat = sum_q g_q M_q,   g = pinv(R_hP_est) hP
// -- N x T taps per column instead of nnz(W) x P weights.  The interference (D-hat - diag h-hat) v of DS.m:482-484 is then
// the same Modulation -> banded channel -> Demodulation chain as the perfect-CSI pass, applied with H-hat.  This equals the
// reference's D-hat up to the entries its two 1e-8 thresholds (DS.m:263-264, 287-289) removed: exactly (1e-13) for a waveform
// where the thresholds only remove rounding noise (CP-OFDM), within 4e-5 of max|D-hat| for FBMC at the default geometry
// (the library measures the removed magnitudes at setup: chest_estimator_info).
//
// k_est_channel: one CTA per EST unit (scheme, SNR point, 16 realizations): g = Rinv hP for the 16 columns, then
// H-hat[column][tap][n] = sum_q g[q][column] M[q][tap][n]; the pseudo-channels are read once per 16 columns.
struct EstChanParams {
    const int4* desc;                                            // per unit slot: {unit, scheme, SNR point, first realization}
    const cplx* hP[3]; const cplx* rinv[3]; int P[3];            // per scheme: pilot estimates [snr][rep][P], Rinv [snr][P x P col-major]
    const cplx* Mq;                                              // [P][T * N] pseudo-channel taps of the waveform
    cplx* hest;                                                  // [unit slot * 16 + column][T * N]
    int n_rep, TN;
};
#define EST_CHAN_THREADS 256
__global__ void __launch_bounds__(EST_CHAN_THREADS) k_est_channel(EstChanParams p) {
    extern __shared__ __align__(16) cplx ec_smem[];
    constexpr int NC = NC_MAX;
    const int4 dsc = p.desc[blockIdx.x];
    IcCta cta; cta.scheme_or_wf = dsc.y; cta.snr = dsc.z; cta.first = dsc.w;
    const int si = cta.scheme_or_wf, P = p.P[si], tid = threadIdx.x, nthr = blockDim.x;
    cplx* hp = ec_smem;                                          // [P][NC]
    cplx* g = hp + P * NC;                                       // [P][NC]
    for (int idx = tid; idx < P * NC; idx += nthr) {
        const int c = idx % NC, pp = idx / NC, rep = cta.first + c;
        hp[idx] = rep < p.n_rep ? p.hP[si][((int64_t)cta.snr * p.n_rep + rep) * P + pp] : cmake(0.0, 0.0);
    }
    __syncthreads();
    const cplx* ri = p.rinv[si] + (int64_t)cta.snr * P * P;
    for (int idx = tid; idx < P * NC; idx += nthr) {
        const int c = idx % NC, q = idx / NC;
        cplx acc = cmake(0.0, 0.0);
        for (int pp = 0; pp < P; ++pp) cfma(acc, ri[q + (int64_t)P * pp], hp[pp * NC + c]);
        g[idx] = acc;
    }
    __syncthreads();
    cplx* out = p.hest + (int64_t)blockIdx.x * NC * p.TN;
    for (int e = tid; e < p.TN; e += nthr) {
        cplx acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = cmake(0.0, 0.0);
        for (int q0 = 0; q0 < P; q0 += 8) {                       // eight pseudo-channel values in flight per thread
            cplx m[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) m[u] = q0 + u < P ? ld_nc(p.Mq + (int64_t)(q0 + u) * p.TN + e) : cmake(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (q0 + u >= P) break;
#pragma unroll
                for (int c = 0; c < NC; ++c) cfma(acc[c], m[u], g[(q0 + u) * NC + c]);
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) out[(int64_t)c * p.TN + e] = acc[c];
    }
}

// The same product on the FP64 tensor pipe (P <= 16): H-hat[e, c] = sum_q M[q][e] g[q][c] as DMMA tiles of 8 samples x 8 columns x 4
// pilots; the g operand stays in registers for the whole unit, the pseudo-channel fragments of the next tile are loaded while the
// current one is multiplied.  Four real products per complex one: g = pinv(R) hP has a large dynamic range at high SNR (the sum
// cancels), and the three-multiplication form measurably costs accuracy here (2e-10 instead of a few 1e-11 on the symbol estimates).
template <int P4>
__global__ void __launch_bounds__(EST_CHAN_THREADS, 2) k_est_channel_mma(EstChanParams p) {
    extern __shared__ __align__(16) cplx ec_smem[];
    constexpr int NC = NC_MAX;
    const int4 dsc = p.desc[blockIdx.x];
    const int si = dsc.y, snr = dsc.z, first = dsc.w, P = p.P[si], tid = threadIdx.x, nthr = blockDim.x;
    cplx* hp = ec_smem;                                          // [P][NC]
    cplx* g = hp + P * NC;                                       // [4 P4][NC], rows >= P zero
    for (int idx = tid; idx < P * NC; idx += nthr) {
        const int c = idx % NC, pp = idx / NC, rep = first + c;
        hp[idx] = rep < p.n_rep ? p.hP[si][((int64_t)snr * p.n_rep + rep) * P + pp] : cmake(0.0, 0.0);
    }
    __syncthreads();
    const cplx* ri = p.rinv[si] + (int64_t)snr * P * P;
    for (int idx = tid; idx < 4 * P4 * NC; idx += nthr) {
        const int c = idx % NC, q = idx / NC;
        cplx acc = cmake(0.0, 0.0);
        if (q < P) for (int pp = 0; pp < P; ++pp) cfma(acc, ri[q + (int64_t)P * pp], hp[pp * NC + c]);
        g[idx] = acc;
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5, gl = lane >> 2, t = lane & 3;
    double b_r[P4][2], b_i[P4][2];                               // B[t][gl] per k-step and n-tile
#pragma unroll
    for (int ks = 0; ks < P4; ++ks)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const cplx b = g[(4 * ks + t) * NC + nt * 8 + gl];
            b_r[ks][nt] = b.x; b_i[ks][nt] = b.y;
        }
    const int TN = p.TN, n_tile = (TN + 7) >> 3;
    cplx* out = p.hest + (int64_t)blockIdx.x * NC * TN;
    cplx a[P4], an[P4];
    auto load = [&](cplx (&dst)[P4], int tile) {
        const int e = tile * 8 + gl;
#pragma unroll
        for (int ks = 0; ks < P4; ++ks) {
            const int q = 4 * ks + t;
            dst[ks] = (tile < n_tile && e < TN && q < P) ? ld_nc(p.Mq + (int64_t)q * TN + e) : cmake(0.0, 0.0);
        }
    };
    load(a, warp);
    for (int tile = warp; tile < n_tile; tile += nwarp) {
        load(an, tile + nwarp);
        double cr[2][2] = {{0, 0}, {0, 0}}, ci[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
        for (int ks = 0; ks < P4; ++ks) {
            const double nai = dneg(a[ks].y);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                dmma884(cr[nt][0], cr[nt][1], a[ks].x, b_r[ks][nt]);
                dmma884(cr[nt][0], cr[nt][1], nai, b_i[ks][nt]);
                dmma884(ci[nt][0], ci[nt][1], a[ks].x, b_i[ks][nt]);
                dmma884(ci[nt][0], ci[nt][1], a[ks].y, b_r[ks][nt]);
            }
        }
        const int e = tile * 8 + gl;
        if (e < TN) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    out[(int64_t)(nt * 8 + 2 * t + j) * TN + e] = cmake(cr[nt][j], ci[nt][j]);
        }
#pragma unroll
        for (int ks = 0; ks < P4; ++ks) a[ks] = an[ks];
    }
}

// k_est_factored: one column (scheme, SNR point, realization) of an EST unit per CTA:
//     y_ic = y - Demodulation(H-hat Modulation(v)) + h-hat v        (h-hat = diag(D-hat) = W_diag hP, left by k_ic_light)
// v from the unit's scratch, y_ic into it (what k_ic_main writes for the tile form of W); the chain is k_perfect_fbmc_det's.
struct EstFactParams {
    ModemDev md;
    const int4* desc;                                           // per unit slot: {unit, scheme, SNR point, first realization}
    const cplx* y[3]; int K_max, n_rep, T, N, K;
    const cplx* hest; const int* tap_delay;
    cplx* scratch;
    int colmajor;                  // unit scratch layout: 1 = [column][K_max], 0 = [row][NC_MAX]
};
template <bool FAST24>
__global__ void __launch_bounds__(PERF_FBMC_THREADS, FAST24 ? CHAIN24_EST_MIN_CTAS : PERF_FBMC_MIN_CTAS) k_est_factored(EstFactParams p) {
    extern __shared__ __align__(16) cplx pf_smem[];
    const ModemDev& md = p.md;
    const int n = md.nfft, Ksym = md.Ksym, L = md.L, N = p.N, K = p.K, TS = md.time_spacing, nx = Ksym * n;
    const int nbuf = max(nx, N);
    const bool fbmc = md.kind == 0;
    const int tid = threadIdx.x, nthr = blockDim.x, col = blockIdx.x, c = col % NC_MAX;
    const int4 dsc = p.desc[col / NC_MAX];                     // {unit, scheme, SNR point, first realization}: one load, not a chain
    const int unit = dsc.x;
    IcCta cta; cta.mode = 0; cta.scheme_or_wf = dsc.y; cta.snr = dsc.z; cta.first = dsc.w; cta.n_cols = NC_MAX;
    const int rep = cta.first + c;
    if (rep >= p.n_rep) return;
    cplx* X0 = pf_smem;
    cplx* X1 = X0 + nbuf;
    cplx* tw = X1 + nbuf;
    double* filt = reinterpret_cast<double*>(tw + n);
    int* bins = reinterpret_cast<int*>(filt + md.Np);
    const cplx* ycol = p.y[cta.scheme_or_wf] + ((int64_t)cta.snr * p.n_rep + rep) * K;
    const int cs = p.colmajor ? 1 : NC_MAX;                   // distance of the column's consecutive rows
    cplx* ub = p.scratch + (int64_t)unit * 3 * p.K_max * NC_MAX + (p.colmajor ? (int64_t)c * p.K_max : c);
    const cplx* hcol = ub;
    const cplx* vcol = ub + (int64_t)p.K_max * NC_MAX;
    cplx* ocol = ub + (int64_t)2 * p.K_max * NC_MAX;
    if (FAST24) {                                              // K <= CHAIN24_NE * 128, N <= CHAIN24_NH * 128 (the host checks)
        if (tid < 24) tw[tid] = md.tw[tid];
        if (fbmc) for (int m = tid; m < md.Np; m += nthr) filt[m] = md.filt[m];
        if (tid < L) bins[tid] = md.bin[tid];
        if (L < 24) { for (int idx = tid; idx < nx; idx += nthr) X0[idx] = cmake(0.0, 0.0); __syncthreads(); }
        cplx e[CHAIN24_NE];
        chain24_head_regs<CHAIN24_NE>(md, X0, vcol, cs, hcol, cs, ycol, K, e);
        __syncthreads();
        const cplx* Y = modem_chain24<CHAIN24_NH>(md, X0, X1, tw, filt, p.hest + (int64_t)col * p.T * N, p.T, p.tap_delay, N);
        chain24_tail_regs<CHAIN24_NE, false>(md, Y, bins, e, nullptr, ocol, cs, K);
        return;
    }
    for (int m = tid; m < n; m += nthr) tw[m] = md.tw[m];
    if (fbmc) for (int m = tid; m < md.Np; m += nthr) filt[m] = md.filt[m];
    for (int m = tid; m < L; m += nthr) bins[m] = md.bin[m];
    if (L < n) for (int idx = tid; idx < nx; idx += nthr) X0[idx] = cmake(0.0, 0.0);
    __syncthreads();
    for (int i = tid; i < K; i += nthr) {
        const int k = i / L, l = i - k * L;
        cplx v = vcol[(int64_t)i * cs];
        if (fbmc) v = cmul(v, md.phase[i]);
        X0[k * n + bins[l]] = cmake(v.x * md.norm, v.y * md.norm);
    }
    __syncthreads();
    const cplx* Y;
    {
        cplx* Xz = fft_shared_batch(X0, X1, tw, md.plan, true, Ksym);
        cplx* Xo = (Xz == X0) ? X1 : X0;
        const double inv_n = 1.0 / n;
        for (int nn = tid; nn < N; nn += nthr) {
            cplx acc = cmake(0.0, 0.0);
            if (fbmc) {                                             // overlap-add (FBMC.m:267-268)
                int k_lo = (nn - md.Np + TS) / TS; if (nn - md.Np + 1 <= 0) k_lo = 0;
                const int k_hi = min(Ksym - 1, nn / TS);
                int tap = nn - k_lo * TS, mm = tap % n;
                const cplx* zc = Xz + k_lo * n;
                for (int k = k_lo; k <= k_hi; ++k) {
                    if (tap >= 0 && tap < md.Np) {
                        const cplx z = zc[mm];
                        const double pf = filt[tap];
                        acc.x = fma(pf, z.x, acc.x); acc.y = fma(pf, z.y, acc.y);
                    }
                    tap -= TS; mm -= TS; if (mm < 0) mm += n;
                    zc += n;
                }
            } else {                                                // cyclic prefix + zero guards (OFDM.m:158-164)
                const int q = nn - md.zero_guard, k = q >= 0 ? q / TS : Ksym;
                if (k < Ksym) { int m = q - k * TS - md.cp; if (m < 0) m += n; acc = Xz[k * n + m]; }
            }
            Xo[nn] = cmake(acc.x * inv_n, acc.y * inv_n);
        }
        __syncthreads();
        const cplx* hr = p.hest + (int64_t)col * p.T * N;
        for (int nn = tid; nn < N; nn += nthr) {                    // r = H-hat s
            cplx acc = cmake(0.0, 0.0);
            for (int t = 0; t < p.T; ++t) {
                const int d = p.tap_delay[t];
                if (nn >= d) cfma(acc, ld_nc(hr + (int64_t)t * N + nn), Xo[nn - d]);
            }
            Xz[nn] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < nx; idx += nthr) {                // fold by O (FBMC.m:297-302) / drop the cyclic prefix (OFDM.m:172-176)
            const int k = idx / n, m = idx - k * n;
            cplx acc = cmake(0.0, 0.0);
            if (fbmc) {
                const cplx* seg = Xz + k * TS + m;
                for (int o = 0; o < md.O; ++o) {
                    const double pf = filt[o * n + m];
                    const cplx v = seg[o * n];
                    acc.x = fma(pf, v.x, acc.x); acc.y = fma(pf, v.y, acc.y);
                }
            } else acc = Xz[md.zero_guard + k * TS + md.cp + m];
            Xo[idx] = acc;
        }
        __syncthreads();
        Y = fft_shared_batch(Xo, Xz, tw, md.plan, false, Ksym);
    }
    for (int i = tid; i < K; i += nthr) {                       // y_ic = y - U + h-hat v
        const int k = i / L, l = i - k * L;
        cplx u0 = Y[k * n + bins[l]];
        if (fbmc) u0 = cmulc(md.phase[i], u0);
        const cplx yv = ycol[i];
        const cplx hvv = cmul(hcol[(int64_t)i * cs], vcol[(int64_t)i * cs]);
        ocol[(int64_t)i * cs] = cmake(yv.x - u0.x * md.inv_demod + hvv.x, yv.y - u0.y * md.inv_demod + hvv.y);
    }
}

// ============================================================================ SimpleVersion_DoublyFlat.m:89-176, batched
// One "body" = one (repetition, SNR point) pass of the script's loop: doubly-flat channel h (one complex scalar), AWGN.
// r[col][n] = h[body] * s[col][n] + sqrt(Pn[body] / 2) * noise[body][wf][n]   (SV.m:123-131), col = g * n_body + body
__global__ void k_sv_channel(cplx* __restrict__ r, const cplx* __restrict__ s, const cplx* __restrict__ h,
                             const cplx* __restrict__ noise, const double* __restrict__ pn, int N, int n_body, int wf) {
    const int n = blockIdx.y * blockDim.x + threadIdx.x, col = blockIdx.x;      // columns on grid.x: up to 2^31 - 1 of them
    if (n >= N) return;
    const int body = col % n_body;
    const double sc = sqrt(pn[body] / 2.0);
    const cplx nz = noise[((int64_t)body * 2 + wf) * N + n];
    cplx v = cmul(h[body], s[(int64_t)col * N + n]);
    r[(int64_t)col * N + n] = cmake(v.x + sc * nz.x, v.y + sc * nz.y);
}
// LS pilot estimates, interpolation h = M hP (M given: K x P, row-major here), one-tap equalisation, de-spreading / selection,
// hard decisions and bit-error counts of one scheme (SV.m:138-169).  One block per body.  err[body][slot] counts; a slot < 0
// is skipped (the script has no perfect-CSI auxiliary branch).
__global__ void k_sv_detect(SchemeDev sd, ConstDev cd, const cplx* __restrict__ y, const cplx* __restrict__ interp,
                            const cplx* __restrict__ htrue, uint32_t* __restrict__ err, int slot_est, int slot_perf, int n_body) {
    extern __shared__ cplx svs[];
    cplx* hP = svs;                    // [P]
    cplx* xe = svs + sd.P;             // [K] equalised with the interpolated channel
    cplx* xp = xe + sd.K;              // [K] equalised with the true channel
    __shared__ unsigned cnt[2];
    const int body = blockIdx.x, tid = threadIdx.x;
    if (body >= n_body) return;
    const cplx* yb = y + (int64_t)body * sd.K;
    if (tid < 2) cnt[tid] = 0;
    for (int p = tid; p < sd.P; p += blockDim.x) {
        const cplx q = cdiv(yb[sd.pilot_pos[p]], sd.xP[(int64_t)body * sd.P + p]);
        hP[p] = cmake(q.x / sd.sqrt_kappa, q.y / sd.sqrt_kappa);
    }
    __syncthreads();
    const cplx ht = htrue[body];
    for (int i = tid; i < sd.K; i += blockDim.x) {
        cplx hh = cmake(0.0, 0.0);
        for (int p = 0; p < sd.P; ++p) cfma(hh, interp[(int64_t)i * sd.P + p], hP[p]);
        xe[i] = cdiv(yb[i], hh);
        xp[i] = cdiv(yb[i], ht);
    }
    __syncthreads();
    unsigned e_est = 0, e_perf = 0, dummy = 0;
    for (int d = tid; d < sd.n_data; d += blockDim.x) {
        cplx a = cmake(0.0, 0.0), b = cmake(0.0, 0.0);
        if (sd.detect_mode == 1) {
            for (int e = sd.ct_colptr[sd.P + d]; e < sd.ct_colptr[sd.P + d + 1]; ++e) {
                const cplx cv = sd.ct_val[e];
                const cplx t1 = cmulc(cv, xe[sd.ct_row[e]]), t2 = cmulc(cv, xp[sd.ct_row[e]]);
                a.x += t1.x; a.y += t1.y; b.x += t2.x; b.y += t2.y;
            }
            a = cmake(a.x / sd.dpr, 0.0); b = cmake(b.x / sd.dpr, 0.0);
        } else {
            const int i = sd.data_pos[d];
            a = cmake(xe[i].x / sd.sqrt_dpr, sd.detect_mode == 0 ? 0.0 : xe[i].y / sd.sqrt_dpr);
            b = cmake(xp[i].x / sd.sqrt_dpr, sd.detect_mode == 0 ? 0.0 : xp[i].y / sd.sqrt_dpr);
        }
        const uint32_t tw = sd.txword[(int64_t)body * sd.n_data + d];
        ic_decide(cd, a, tw, 0u, e_est, dummy);
        ic_decide(cd, b, tw, 0u, e_perf, dummy);
    }
    if (e_est) atomicAdd(&cnt[0], e_est);
    if (e_perf) atomicAdd(&cnt[1], e_perf);
    __syncthreads();
    if (tid == 0) {
        if (slot_est >= 0) err[(int64_t)body * 5 + slot_est] = cnt[0];
        if (slot_perf >= 0) err[(int64_t)body * 5 + slot_perf] = cnt[1];
    }
}

// ============================================================================ setup on the device (DS.m:208-313)
// Pseudo-channel of pilot p (DS.m:213,260): M_p = reshape(R_vecH * kron(g_p.', q_p')', N, N) is banded like H, with
//   M_p[a + m, a] = pdp_m * sum_a' rt[a - a'] * zeta_m[a'],   zeta_m[a'] = q_p[row(a')] * conj(g_p[col(a')])
// where (row, col) = position of entry a'(N+1) + m of H(:) (FF.m:377: regular entries row = a' + m, col = a'; entries
// running past the bottom of a column wrap to (a' + m - N, a' + 1), those beyond N^2 are cropped, FF.m:406).  Regular
// entries go to h[p][tap][a + m] (the layout K1 / K2 read), wrapped ones to corner[p][tap][a + m - N].
__global__ void k_pseudo_channel(cplx* __restrict__ h, cplx* __restrict__ corner, const cplx* __restrict__ G,
                                 const cplx* __restrict__ Q, const int* __restrict__ pil, const int* __restrict__ g_lo,
                                 const int* __restrict__ g_hi, const double* __restrict__ rt, const int* __restrict__ tap_delay,
                                 const double* __restrict__ tap_pow, int N, int T, int max_delay) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x, tap = blockIdx.y, p = blockIdx.z;
    if (a >= N) return;
    const int m = tap_delay[tap], sym = pil[p];
    const cplx* q = Q + (int64_t)N * sym;
    const cplx* g = G + (int64_t)N * sym;
    const double* r0 = rt + (N - 1) + a;                       // rt[(N-1) + a - a']
    cplx acc = cmake(0.0, 0.0);
    const int lo = g_lo[sym], hi = min(g_hi[sym], N - m);       // regular a': a' + m < N, g[a'] != 0
    for (int ap = lo; ap < hi; ++ap) {
        const cplx z = cmulc(g[ap], q[ap + m]);                 // q * conj(g)
        const double w = r0[-ap];
        acc.x = fma(w, z.x, acc.x); acc.y = fma(w, z.y, acc.y);
    }
    for (int ap = max(N - m, 0); ap <= N - 2; ++ap) {           // wrapped a': row = a' + m - N, col = a' + 1
        const cplx z = cmulc(g[ap + 1], q[ap + m - N]);
        const double w = r0[-ap];
        acc.x = fma(w, z.x, acc.x); acc.y = fma(w, z.y, acc.y);
    }
    const double pw = tap_pow[tap];
    acc.x *= pw; acc.y *= pw;
    if (a + m < N) h[((int64_t)p * T + tap) * N + a + m] = acc;
    else if (a <= N - 2) corner[((int64_t)p * T + tap) * max_delay + (a + m - N)] = acc;
    if (a < m) h[((int64_t)p * T + tap) * N + a] = cmake(0.0, 0.0);      // rows above the band start
}

// R[p][e] = D_p[e] + wrapped-entry contributions  v * conj(Q[r, i]) * G[c, j]  (exactness of FF.m:377), then the
// |.| < thr -> 0 rule of DS.m:263-264.  Layout of D and R: row-tile-major [p][K/8][K][8].
__global__ void k_rsup_finish(cplx* __restrict__ R, const cplx* __restrict__ D, const cplx* __restrict__ corner,
                              const cplx* __restrict__ G, const cplx* __restrict__ Q, const int* __restrict__ tap_delay,
                              int N, int K, int T, int max_delay, double thr, unsigned long long* __restrict__ zmax) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    const int RT8 = ((K + 7) / 8) * 8;
    if (e >= (int64_t)RT8 * K) return;
    const int r8 = (int)(e & 7), j = (int)((e >> 3) % K), i = (int)((e >> 3) / K) * 8 + r8;
    cplx v = D[(int64_t)p * RT8 * K + e];
    if (i < K) {
        for (int t = 0; t < T; ++t) {
            const int m = tap_delay[t];
            for (int r = 0; r + 2 <= m; ++r) {
                const cplx cv = corner[((int64_t)p * T + t) * max_delay + r];
                if (cv.x == 0.0 && cv.y == 0.0) continue;
                const cplx qv = Q[(int64_t)N * i + r];
                if (qv.x == 0.0 && qv.y == 0.0) continue;
                const cplx gv = G[(int64_t)N * j + (N - m + r + 1)];
                cfma(v, cmul(cv, cmake(qv.x, -qv.y)), gv);
            }
        }
        const double mag = hypot(v.x, v.y);
        if (mag < thr) {                                        // largest magnitude the threshold removes (bit pattern order = value order)
            if (zmax && mag > 0.0) atomicMax(zmax, (unsigned long long)__double_as_longlong(mag));
            v = cmake(0.0, 0.0);
        }
    } else v = cmake(0.0, 0.0);
    R[(int64_t)p * RT8 * K + e] = v;
}

// R_hP[p'][p] = D_p[pil[p'], pil[p']] (DS.m:213), column-major P x P; before thresholding
__global__ void k_rhp_gather(cplx* __restrict__ out, const cplx* __restrict__ R, const int* __restrict__ pil, int K, int P) {
    const int pp = threadIdx.x, p = blockIdx.x;
    if (pp >= P) return;
    const int RT8 = ((K + 7) / 8) * 8, i = pil[pp];
    out[pp + (int64_t)P * p] = R[(int64_t)p * RT8 * K + ((int64_t)(i >> 3) * K + i) * 8 + (i & 7)];
}

// W[(i,j), p'] = sum_p R_p[(i,j)] Rinv[p, p']  (DS.m:283-313: W = R_Dij_hP * pinv(R_hP_est)), |.| < thr -> 0.
// pass A: mask[e] = 1 where any (snr, p') entry survives.  One thread per (entry e, p'), snr = blockIdx.y.
__global__ void k_w_mask(int* __restrict__ mask, const cplx* __restrict__ R, const cplx* __restrict__ Rinv, int K, int P, double thr,
                         unsigned long long* __restrict__ zmax) {
    const int RT8 = ((K + 7) / 8) * 8;
    const int64_t n_e = (int64_t)RT8 * K;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int pp = (int)(idx % P);
    const int64_t e = idx / P;
    if (e >= n_e) return;
    const cplx* ri = Rinv + (int64_t)blockIdx.y * P * P + (int64_t)P * pp;
    cplx w = cmake(0.0, 0.0);
    bool any = false;
    for (int p = 0; p < P; ++p) {
        const cplx a = R[(int64_t)p * n_e + e];
        if (a.x != 0.0 || a.y != 0.0) { any = true; cfma(w, a, ri[p]); }
    }
    if (!any) return;
    const double mag = hypot(w.x, w.y);
    if (!(mag < thr)) mask[e] = 1;
    else if (zmax && mag > 0.0) atomicMax(zmax, (unsigned long long)__double_as_longlong(mag));
}
// The same pass with one thread per (entry e, SNR point): the P correlations of the entry are loaded once and held in registers,
// pinv(R)[snr] sits in shared memory (P <= 32).  k_w_mask re-reads them once per output column p': 16 x the traffic, which made
// it the largest term of the per-velocity setup of a sweep.
template <int PMAX>
__global__ void k_w_mask_rows(int* __restrict__ mask, const cplx* __restrict__ R, const cplx* __restrict__ Rinv, int K, int P, double thr,
                              unsigned long long* __restrict__ zmax) {
    extern __shared__ __align__(16) cplx wm_smem[];
    const int RT8 = ((K + 7) / 8) * 8;
    const int64_t n_e = (int64_t)RT8 * K;
    const cplx* ri = Rinv + (int64_t)blockIdx.y * P * P;
    for (int idx = threadIdx.x; idx < P * P; idx += blockDim.x) wm_smem[idx] = ri[idx];
    __syncthreads();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double zm = 0.0;
    if (e < n_e) {
        cplx a[PMAX];
        bool any = false;
#pragma unroll
        for (int p = 0; p < PMAX; ++p) {
            a[p] = p < P ? R[(int64_t)p * n_e + e] : cmake(0.0, 0.0);
            any |= (a[p].x != 0.0 || a[p].y != 0.0);
        }
        if (any) {
            bool keep = false;
            for (int pp = 0; pp < P; ++pp) {
                cplx w = cmake(0.0, 0.0);
#pragma unroll
                for (int p = 0; p < PMAX; ++p) if (p < P && (a[p].x != 0.0 || a[p].y != 0.0)) cfma(w, a[p], wm_smem[p + P * pp]);
                const double mag = hypot(w.x, w.y);
                if (!(mag < thr)) keep = true;
                else zm = fmax(zm, mag);
            }
            if (keep) mask[e] = 1;
        }
    }
    if (zmax) {                                                 // one atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) zm = fmax(zm, __shfl_xor_sync(0xffffffffu, zm, o));
        if ((threadIdx.x & 31) == 0 && zm > 0.0) atomicMax(zmax, (unsigned long long)__double_as_longlong(zm));
    }
}
// pass B: fragments of one SNR point.  One thread per (tile t, row r, p'); the diagonal goes to dg / dfrag.
__global__ void k_w_fill(cplx* __restrict__ frag, cplx* __restrict__ dg, cplx* __restrict__ dfrag, const cplx* __restrict__ R,
                         const cplx* __restrict__ Rinv, const int* __restrict__ tile_rt, const int* __restrict__ tile_delta,
                         int n_tiles, int K, int P, int P4, double thr) {
    const int RT8 = ((K + 7) / 8) * 8;
    const int64_t n_e = (int64_t)RT8 * K;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int pp = (int)(idx % P), r = (int)((idx / P) & 7);
    const int64_t t = idx / P / 8;
    if (t >= n_tiles + (K + 7) / 8) return;
    const bool diag = t >= n_tiles;                               // the last RT "tiles" are the diagonal
    const int rt = diag ? (int)(t - n_tiles) : tile_rt[t];
    const int i = rt * 8 + r, j = diag ? i : i + tile_delta[t];
    if (i >= K || j < 0 || j >= K) return;
    const int64_t e = ((int64_t)rt * K + j) * 8 + r;
    const cplx* ri = Rinv + (int64_t)P * pp;
    cplx w = cmake(0.0, 0.0);
    for (int p = 0; p < P; ++p) {
        const cplx a = R[(int64_t)p * n_e + e];
        if (a.x != 0.0 || a.y != 0.0) cfma(w, a, ri[p]);
    }
    if (hypot(w.x, w.y) < thr) w = cmake(0.0, 0.0);
    if (diag) {
        dg[(int64_t)i * P + pp] = w;
        dfrag[((int64_t)rt * P4 + (pp >> 2)) * 32 + r * 4 + (pp & 3)] = w;
    } else frag[((int64_t)t * P4 + (pp >> 2)) * 32 + r * 4 + (pp & 3)] = w;
}

// ============================================================================ counter totals
// tot[e] += sum_rep err[rep][e]  (e = (snr, it, scheme, csi, edge)): the per-GPU partial sums of the final reduce.
// One block per counter; 64-bit totals (4096 realizations x 5504 bits already exceed 2^24 per batch).
__global__ void k_sum_counters(unsigned long long* __restrict__ tot, const uint32_t* __restrict__ err, int n_rep, int per_rep) {
    const int e = blockIdx.x;
    unsigned long long a = 0;
    for (int r = threadIdx.x; r < n_rep; r += blockDim.x) a += err[(int64_t)r * per_rep + e];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __shared__ unsigned long long part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += part[w];
        tot[e] += t;
    }
}

// ============================================================================ FP64 peak probes
__global__ void k_peak_dmma(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
// DMMA and DFMA interleaved: do the two FP64 pipes add up?
__global__ void k_peak_mix(double* out, int iters) {
    double c[4][2], f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-9 + i;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dmma884(c[i][0], c[i][1], a, b);
            f[2 * i] = fma(f[2 * i], a, b);
            f[2 * i + 1] = fma(f[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void k_peak_dfma(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}
