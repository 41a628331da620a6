/*
 * chest_b200.h -- C ABI of the B200-native Monte-Carlo hot path of rnissel/Channel-Estimation.
 *
 * This is the drop-in boundary: a MATLAB MEX gateway (matlab/chest_mex.c), the Python host mirror
 * (channel-estimation_b200/context.py + _lib.py, ctypes) or any other FFI binds exactly these entry points.
 * The reference has no native interface of its own (it is 100 % MATLAB); each entry point cites
 * the reference method / script lines it replaces.  `DS.m` = DoublySelectiveChannelEstimation.m,
 * `FF.m` = +Channel/FastFading.m, `FBMC.m`/`OFDM.m`/`SC.m` live in +Modulation.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure; the message is
 *     available from chest_last_error() (thread-local, valid until the next call);
 *   - all host arrays are column-major (MATLAB order); complex arrays are interleaved
 *     (re,im) double pairs; indices are 0-based;
 *   - a context owns all device memory; nothing here takes or returns torch / gpuArray types;
 *   - there is NO CPU fallback: without an sm_100 device chest_create fails.
 */
#ifndef CHEST_B200_H
#define CHEST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHEST_OK               0
#define CHEST_ERR_ARG         -1
#define CHEST_ERR_CUDA        -2
#define CHEST_ERR_STATE       -3
#define CHEST_ERR_NO_DEVICE   -4

/* schemes simulated side by side on the same channel + noise (DS.m:371-373,399-403) */
#define CHEST_SCHEME_AUX   0   /* FBMC, auxiliary symbols */
#define CHEST_SCHEME_COD   1   /* FBMC, data spreading    */
#define CHEST_SCHEME_OFDM  2
#define CHEST_N_SCHEMES    3
/* waveforms (one G/Q pair each, DS.m:191-195) */
#define CHEST_WF_FBMC      0
#define CHEST_WF_OFDM      1
#define CHEST_N_WF         2
/* how equalised symbols become data-symbol estimates */
#define CHEST_DETECT_SELECT_REAL     0  /* real(x(idx)/sqrt(dpr))        DS.m:430,523 */
#define CHEST_DETECT_DESPREAD_REAL   1  /* real((C'*x)(P+1:end))/dpr     DS.m:436-437,520,524 */
#define CHEST_DETECT_SELECT_COMPLEX  2  /* x(idx)/sqrt(dpr)              DS.m:444,525 */
/* constellations (SC.m) */
#define CHEST_CONST_PAM    0
#define CHEST_CONST_QAM    1
/* Doppler models of Channel.FastFading.NewRealization (FF.m:225-232) */
#define CHEST_DOPPLER_JAKES    0
#define CHEST_DOPPLER_UNIFORM  1
/* 'Discrete-Jakes' / 'Discrete-Uniform' (FF.m:151-182, 203-221): 2 n_shift + 1 discrete Doppler shifts spaced fs/N, the
 * realization is the inverse DFT of their weighted complex normals (a pruned inverse DFT on the device) */
#define CHEST_DOPPLER_DISCRETE_JAKES    2
#define CHEST_DOPPLER_DISCRETE_UNIFORM  3
/* MMSE matrix variants (DS.m:492): W_MMSE / W_MMSE_noInterference */
#define CHEST_W_WITH_INTERFERENCE  0
#define CHEST_W_NO_INTERFERENCE    1

const char* chest_last_error(void);

/* Library / device probe.  n_sm, cc_major, cc_minor may be NULL. */
int chest_device_info(int device, int* n_sm, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ context life-cycle */
int chest_create(int device, uint64_t* handle);
int chest_destroy(uint64_t handle);

/* ------------------------------------------------------------------ one-time setup
 * (the outputs of DS.m:50-313, which stay on the host)            */

/* Channel.FastFading constructor state (FF.m:25-192): N samples, the normalised power delay
 * profile over ALL Lt taps (zeros allowed, FF.m:129), maximum Doppler shift, dt, paths.
 * max_doppler_hz = 0 declares a time-invariant channel (block fading FF.m:241-248, AWGN FF.m:197-198):
 * realizations are then uploaded with chest_set_impulse_response and applied by the same banded operator;
 * chest_new_realization* and chest_run_batch require a positive Doppler shift. */
int chest_set_channel(uint64_t handle, int n_samples, int n_taps, const double* pdp_normalized,
                      double max_doppler_hz, double dt, int n_paths, int doppler_model);

/* G = GetTXMatrix (N x K) and Q = GetRXMatrix' (N x K) of one waveform
 * (FBMC.m:318-354, OFDM.m:184-218; DS.m:191-195).  Complex interleaved, column-major. */
int chest_set_waveform(uint64_t handle, int waveform, int n_samples, int n_symbols_k,
                       const double* G, const double* Q);

/* SignalConstellation tables (SC.m:24-74): SymbolMapping (order complex values, sorted by bit
 * value) and BitMapping (order x nbits, column-major, 0/1 bytes). */
int chest_set_constellation(uint64_t handle, int which, int order,
                            const double* symbol_mapping, const uint8_t* bit_mapping);

/* One transmission scheme: precoder C (K x K_in sparse CSC, complex; AuxiliaryMethod /
 * CodingMethod .PrecodingMatrix or PilotMapping_OFDM, DS.m:116-137), pilot positions
 * (find(PilotMatrix==1)), data positions (rows selected at DS.m:430/444; NULL for the
 * de-spreading scheme), kappa (DS.m:140-142), data power reduction, detection mode, and the
 * "NoEdge" bit mask ConsideredBits_* (DS.m:170-172; n_data*nbits bytes). */
int chest_set_scheme(uint64_t handle, int scheme, int waveform, int k_in, int n_pilots, int n_data,
                     const int64_t* c_jc, const int32_t* c_ir, const double* c_val,
                     const int32_t* pilot_pos, const int32_t* data_pos,
                     double kappa, double data_power_reduction, int detect_mode, int constellation,
                     const uint8_t* considered_bits);

/* Noise powers Pn_time per SNR point (DS.m:398). */
int chest_set_snr(uint64_t handle, int n_snr, const double* pn_time);

/* W_MMSE_* / W_MMSE_noInterference_* of one scheme (DS.m:279-313): a (K^2 P) x n_snr sparse
 * matrix in CSC form exactly as MATLAB stores it (jc: n_snr+1, ir: row = i + K j + K^2 p). */
int chest_set_mmse(uint64_t handle, int scheme, int variant, int n_snr,
                   const int64_t* jc, const int64_t* ir, const double* val);

/* The same setup on the device (DS.m:208-313).  chest_setup_correlations needs a context finalized for at least
 * n_pilots realizations: it synthesises the banded pseudo-channels M_p = reshape(R_vecH*kron(g_p.',q_p')',N,N) of all
 * pilots from the time correlation (FF.m:321-340: 2N-1 real values, lag 0 at index N-1) and the power delay profile
 * (FF.m:366-407, including the wrapped entries of FF.m:377), runs them through K1 + K2 (D_p = Q' M_p G, DS.m:260), returns
 * R_hP = D_p(pilot, pilot) (DS.m:213; P x P complex, column-major) and keeps R_Dij_hP -- with the zero_threshold rule of
 * DS.m:263-264 applied -- on the device.  n_support (may be NULL): number of (i,j) with a non-zero row of R_Dij_hP.
 * chest_build_mmse then forms W = R_Dij_hP * R_inv for every SNR point (R_inv: n_snr matrices pinv(R_hP_est), P x P complex
 * column-major, DS.m:283-285: P x P host work), applies the second threshold (DS.m:287-289) and writes the diagonal-tile
 * format directly -- the (K^2 P) x n_snr sparse matrix never exists on the host.  It replaces chest_set_mmse.
 * chest_release_setup frees R_Dij_hP (P x K^2 complex) once every scheme has its matrices. */
int chest_setup_correlations(uint64_t handle, int waveform, int n_pilots, const int32_t* pilot_pos,
                             const double* time_correlation, double zero_threshold, double* R_hP_out, int64_t* n_support);
int chest_build_mmse(uint64_t handle, int scheme, int variant, int n_snr, const double* R_inv, double zero_threshold);
int chest_release_setup(uint64_t handle);

/* Allocate device state for batches of up to max_batch realizations.  Must follow the setters. */
int chest_finalize(uint64_t handle, int max_batch);

/* ------------------------------------------------------------------ tier 1: method-level calls */

/* Channel.FastFading.NewRealization (FF.m:194-250, Jakes/Uniform branch) for `batch` independent
 * realizations.  doppler_u / phase_u: batch x (T x paths) uniforms in the order MATLAB's
 * rand([T 1 paths]) produces them (tap fastest); host pointers. */
int chest_new_realization(uint64_t handle, int batch, const double* doppler_u, const double* phase_u);
/* Same, drawing the uniforms on the device with the counter-based generator
 * (seed, first_rep + b) -- see DESIGN.md "Random numbers". */
int chest_new_realization_seeded(uint64_t handle, int batch, uint64_t seed, int64_t first_rep);

/* NewRealization of a 'Discrete-*' channel from explicit normals (layout of chest_draws.channel_gauss). */
int chest_new_realization_gauss(uint64_t handle, int batch, const double* gauss);
/* What the channel constructor derived: number of discrete Doppler shifts per side (0 unless 'Discrete-*'), the maximum
 * Doppler shift in use (FF.m:146-149 sets it to 0 when too low for a discrete spectrum), non-zero taps.  Any may be NULL. */
int chest_channel_info(uint64_t handle, int* n_doppler_shifts, double* max_doppler_hz, int* n_nonzero_taps);

/* Load `batch` time-variant impulse responses computed elsewhere (channel realizations exported
 * from the reference, or the banded pseudo-channels of the correlation setup, DS.m:213,260):
 * h is batch x (N x Lt) complex, column-major per realization; only the non-zero taps are used. */
int chest_set_impulse_response(uint64_t handle, int batch, const double* h);
/* obj.ImpulseResponse(:,:,1,1) of realization b: N x Lt complex (FF.m:237). */
int chest_get_impulse_response(uint64_t handle, int b, double* out);
/* GetConvolutionMatrix{1,1} of realization b as CSC (FF.m:276-295): jc[N+1], ir[nnz], val[nnz]
 * complex; nnz = sum over non-zero taps m of (N - m).  Pass NULLs to query nnz only. */
int chest_get_convolution_csc(uint64_t handle, int b, int64_t* nnz, int64_t* jc, int32_t* ir, double* val);
/* Convolution(signal) / ConvolutionMatrix*s (FF.m:253-264, DS.m:383-385): r = H_b * s for
 * n_cols column vectors of length N (host, complex). */
int chest_convolve(uint64_t handle, int b, const double* s, int n_cols, double* r);

/* D = Q'*H_b*G and h = diag(D) (DS.m:388-393) of realization b; D_out K x K, h_out K (may be NULL). */
int chest_transmission_matrix(uint64_t handle, int b, int waveform, double* D_out, double* h_out);

/* The same product for realizations 0 .. n_rep-1 of the current channel batch in one go, D left on the device (row-tile-major;
 * the structural zeros are never touched): the "chunked D GEMM" of the scaled-bandwidth workloads, where one D is 1.3 GB
 * (K = 9000).  ms (may be NULL): device time of the banded apply H*G (k_apply_hg) and of the GEMM (k_gemm_d);
 * flops_per_realization (may be NULL): support-aware algorithmic flops; h_out (may be NULL): diag(D) of the
 * n_rep realizations, n_rep x K complex (host pointer; DS.m:391-393).  chest_transmission_matrix_entries
 * reads n entries D[rows[e], cols[e]] of realization b back (complex interleaved) without moving the whole matrix. */
int chest_transmission_matrix_batch(uint64_t handle, int n_rep, int waveform, float* ms /* [2] */, double* flops_per_realization,
                                    double* h_out);
int chest_transmission_matrix_entries(uint64_t handle, int b, int waveform, int n, const int32_t* rows, const int32_t* cols, double* out);

/* s = G*x (Modulation, FBMC.m:319-320 / OFDM.m:185-186) and y = Q'*r (Demodulation,
 * FBMC.m:344-345 / OFDM.m:206-207) in their matrix form for n_cols columns. */
int chest_modulate(uint64_t handle, int waveform, const double* x, int n_cols, double* s);
int chest_demodulate(uint64_t handle, int waveform, const double* r, int n_cols, double* y);

/* The same two methods in their FFT form (FBMC.m:255-268, 287-302 polyphase branch; OFDM.m:153-181), for waveforms whose
 * dense G / Q would not fit or would be wasteful (SimpleVersion_DoublyFlat.m:118-135, scaled bandwidths).  chest_set_modem
 * takes what SetDependentParameters computes (FBMC.m:61-160, OFDM.m:53-88): kind 0 = FBMC polyphase, 1 = CP-OFDM;
 * bin_of_subcarrier[l] = FFT bin (0-based) that carries subcarrier l (rows of IndexPolyphaseMap in ascending order /
 * IntermediateFrequency + l); FBMC: time_spacing = FFTSize/2, overlapping factor O (prototype filter: O*FFTSize real
 * taps), PhaseShift L x K complex, NormalizationFactor, SubcarrierSpacing; OFDM: cyclic prefix and zero-guard samples,
 * time_spacing = FFTSize + CP.  The transforms are hand-written mixed-radix shared-memory FFTs (prime factors <= 13).
 * x: (L*K) x n_cols, s / r: N x n_cols, y: (L*K) x n_cols; complex, column-major, host pointers. */
int chest_set_modem(uint64_t handle, int waveform, int kind, int n_subcarriers, int n_mc_symbols, int fft_size,
                    const int32_t* bin_of_subcarrier, int time_spacing, int overlapping_factor, int cyclic_prefix,
                    int zero_guard_samples, const double* prototype_filter, const double* phase_shift,
                    double normalization_factor, double subcarrier_spacing);
int chest_modulate_fft(uint64_t handle, int waveform, const double* x, int n_cols, double* s);
int chest_demodulate_fft(uint64_t handle, int waveform, const double* r, int n_cols, double* y);

/* ------------------------------------------------------------------ SimpleVersion_DoublyFlat.m:89-176 as a batched launch
 * One "body" = one (repetition, SNR point) pass of the script's double loop: bits -> symbols -> precoding (SV.m:95-115),
 * FFT-form Modulation (SV.m:118-120), doubly-flat channel h ~ CN(0,1) plus noise of power Pn_time (SV.m:123-131),
 * Demodulation (SV.m:133-135), LS pilot estimates (SV.m:138-140), interpolation (SV.m:143-145: the interpolation MATRIX
 * of PilotSymbolAidedChannelEstimation.GetInterpolationMatrix, PSACE.m:171-184, is an input -- MATLAB's scatteredInterpolant
 * is closed source), one-tap equalisation, de-spreading / selection, hard decisions and bit-error counts (SV.m:148-169).
 * The context needs chest_set_modem for both waveforms (no dense G / Q, no channel), the constellations and the three
 * schemes (chest_set_scheme; kappa = the LS scaling of SV.m:138-140, dpr = AuxiliaryMethod.DataPowerReduction for the
 * auxiliary scheme and 1 otherwise), chest_set_interpolation per scheme (K x P complex, column-major) and chest_finalize.
 * err_out[body][5]: bit errors of FBMC-Aux, FBMC-Cod, FBMC perfect CSI, OFDM, OFDM perfect CSI (BER_* of SV.m:165-169 =
 * err / n_bits).  draws == NULL: counter-based generator keyed by (seed, first_body + b). */
typedef struct chest_sv_draws {
    const uint8_t* bits[CHEST_N_SCHEMES];  /* n_body x n_bits(scheme)                       SV.m:95-97   */
    const int32_t* pilot_idx[CHEST_N_WF];  /* n_body x P, 0-based                            SV.m:105,107 */
    const double*  h;                      /* n_body complex: the channel h = sqrt(1/2)*(randn+1j*randn) itself  SV.m:123 */
    const double*  noise[CHEST_N_WF];      /* per waveform n_body x N complex standard normals SV.m:125-126 */
    int            on_device;              /* must be 0 */
} chest_sv_draws;
int chest_set_interpolation(uint64_t handle, int scheme, const double* interpolation_matrix);
int chest_sv_run_batch(uint64_t handle, int n_body, const double* pn_time, const chest_sv_draws* draws, uint64_t seed,
                       int64_t first_body, uint32_t* err_out);

/* D_est = sum_p W(:,:,p) hP(p), h_est = diag(D_est) (DS.m:417-428,493-517).
 * hP: P complex; Dhat_out K x K (may be NULL); hdiag_out K (may be NULL). */
int chest_estimate(uint64_t handle, int scheme, int variant, int i_snr, const double* hP,
                   double* Dhat_out, double* hdiag_out);

/* ------------------------------------------------------------------ tier 2: the loop body */

/* Explicit random draws of n_rep realizations, in the order DS.m:352-368,399 consumes them.
 * With on_device != 0 the pointers are device pointers (inputs resident in HBM). */
typedef struct chest_draws {
    const double*  doppler_u;      /* n_rep x (T x paths)                          FF.m:227 */
    const double*  phase_u;        /* n_rep x (T x paths)                          FF.m:233 */
    const uint8_t* bits[CHEST_N_SCHEMES];  /* n_rep x n_bits(scheme), 0/1         DS.m:355-357 */
    const int32_t* pilot_idx[CHEST_N_WF];  /* n_rep x P, 0-based SymbolMapping idx DS.m:365,367 */
    const double*  noise;          /* n_rep x n_snr x N complex, randn+1j*randn    DS.m:399 */
    int            on_device;
    /* 'Discrete-*' Doppler models only (doppler_u / phase_u are then unused): n_rep x (2 n_shift + 1) x T complex standard
     * normals, rows 0..n_shift = GaussUncorr1 / sqrt(N^2/2), rows n_shift+1..2 n_shift = GaussUncorr2   FF.m:208-209 */
    const double*  channel_gauss;
} chest_draws;

/* Asynchronous upload of host draws into one of two library-owned device buffer sets (a dedicated copy
 * stream): returns at once with `dev` pointing at the device copies (on_device = 1).  The chest_run_batch*
 * call that receives `dev` waits for the copy; the draws of batch i+1 can therefore travel while batch i runs
 * (call order: prefetch(i+1), run(i)).  A set is reused by every second call, after the batch that read it. */
int chest_prefetch_draws(uint64_t handle, int n_rep, const chest_draws* host, chest_draws* dev);

/* Bytes a host->device upload of `n_rep` realizations' draws moves (for reporting). */
int64_t chest_draws_bytes(uint64_t handle, int n_rep);

/* DS.m:350-565 for n_rep realizations: new channel, TX, true D / h, and for every SNR point the
 * MMSE estimate, one-tap equaliser, n_iter interference-cancellation iterations, the perfect-CSI
 * twin and all bit-error counts.  Output: err[rep][snr][it][scheme][csi][edge] uint32 counts,
 * it = 0 (one-tap) .. n_iter, csi 0 = estimated / 1 = perfect, edge 0 = all bits / 1 = NoEdge.
 * BER_* of DS.m:322-345 = err / n_bits (chest_bit_counts).  err_out is a host pointer.
 * Either draws (explicit) or, if draws == NULL, the device generator keyed by (seed, first_rep+r). */
int chest_run_batch(uint64_t handle, int n_rep, int n_iter, const chest_draws* draws,
                    uint64_t seed, int64_t first_rep, uint32_t* err_out);
/* How the perfect-CSI twin of the loop body applies the true channel (DS.m:541-543, y - (D - diag h) v):
 * CHEST_PERFECT_FACTORED (default): y - Q^H (H (G v)) + h v with h = diag(D) from a small GEMM -- D itself is never
 * formed (exactly as D-hat never is);  CHEST_PERFECT_DENSE: D = Q^H H G is materialised per realization (K2) and
 * applied.  Same results; chest_transmission_matrix returns D in either mode. */
#define CHEST_PERFECT_DENSE     0
#define CHEST_PERFECT_FACTORED  1
int chest_set_perfect_csi_mode(uint64_t handle, int mode);
/* Arithmetic of the estimated-CSI interference cancellation y - (D_est - diag h_est) v (DS.m:482-484, the dominant cost of
 * the loop body).  CHEST_PRECISION_FP64 (default): FP64 tensor-core DMMA, results within 1e-9 of the reference arithmetic and
 * identical hard decisions.  CHEST_PRECISION_SPLIT_BF16: the stated reduced-precision mode -- the same sum on the 5th-generation
 * tensor cores (tcgen05.mma, FP32 accumulators in tensor memory, both operands split into two BF16 slices, three products):
 * the cancelled symbols differ by ~1e-5 of the interference magnitude (within the 1e-4 the mode states); pilot estimates,
 * W_diag h_P, equalisation, decisions, counters and the perfect-CSI twin stay FP64.  Needs CHEST_PERFECT_FACTORED and at most 32
 * pilots per scheme; the operand images are packed at the next run.  chest_precision_info: mode, the dense BF16 flops one
 * launch of the tensor-core kernel executes (roofline numerator) and the bytes of the packed operand images; any may be NULL. */
#define CHEST_PRECISION_FP64        0
#define CHEST_PRECISION_SPLIT_BF16  1
int chest_set_precision(uint64_t handle, int mode);
int chest_precision_info(uint64_t handle, int* mode, double* mma_flops_per_launch, int64_t* operand_bytes);
/* Form of the estimated-CSI interference cancellation.  The MMSE estimate is linear in the pilot estimates,
 *     D_est = sum_p W_p hP(p),  W = R_Dij_hP pinv(R_hP_est),  R_Dij_hP(:, q) = vec(Q' M_q G)           (DS.m:256-289, 417-425)
 * with M_q the banded pseudo-channel of pilot q (DS.m:260), hence D_est = Q' H_est G with the estimated channel
 * H_est = sum_q g_q M_q, g = pinv(R_hP_est) hP: N x taps values per column instead of nnz(W) x P weights, and
 * (D_est - diag h_est) v becomes Modulation -> banded channel -> Demodulation with H_est (FFT modem of chest_set_modem), ~25 x
 * fewer operations than the weights at the default geometry.  It equals the reference's D_est up to what its two 1e-8
 * thresholds (DS.m:263-264, 287-289) removed from R_Dij_hP and W; the device-side setup records the largest removed magnitudes.
 *   CHEST_ESTIMATOR_AUTO (default)  the factored form only for a scheme where the thresholds removed nothing above 1e-13
 *                                   (CP-OFDM: only rounding noise is removed, results within 1e-12 of the tile form of W) AND
 *                                   the tile form costs more than 3 MFLOP per column (the measured break-even of the modem
 *                                   chain; CP-OFDM at the default geometry stays on the tiles); the thresholded W tiles
 *                                   otherwise (bit-faithful to the reference's W)
 *   CHEST_ESTIMATOR_FACTORED_EXACT  the factored form wherever the thresholds removed nothing above 1e-13, whatever it costs
 *   CHEST_ESTIMATOR_TILES           the thresholded W for every scheme
 *   CHEST_ESTIMATOR_FACTORED        the factored form for every scheme that has the factors: a STATED-TOLERANCE mode -- D_est
 *                                   differs from the reference's by the removed entries (4e-5 of max|D_est| for FBMC at the
 *                                   default geometry, within the 1e-4 of the stated reduced-accuracy mode); FP64 arithmetic
 * Needs CHEST_PERFECT_FACTORED, a modem description that reproduces G / Q (chest_set_modem), pseudo-channels without wrapped
 * entries (tap delays < 2 samples) and the factors: kept automatically by chest_setup_correlations + chest_build_mmse, or
 * uploaded -- chest_set_pseudo_channels: M[p][tap][n] = M_p(n, n - delay_tap) (taps = non-zero PDP entries, ascending delay),
 * chest_set_estimator_factors: R_inv as in chest_build_mmse; removed_max = largest magnitude the caller's thresholds removed
 * (a value below 1e-13 lets CHEST_ESTIMATOR_AUTO pick the factored form).  chest_estimator_info (after a run; any pointer may be
 * NULL): mode, whether the scheme ran factored, the removed magnitudes, kernel time of the factored pass in the last profiled run. */
#define CHEST_ESTIMATOR_AUTO      0
#define CHEST_ESTIMATOR_TILES     1
#define CHEST_ESTIMATOR_FACTORED  2
#define CHEST_ESTIMATOR_FACTORED_EXACT 3
int chest_set_estimator_mode(uint64_t handle, int mode);
int chest_set_pseudo_channels(uint64_t handle, int waveform, int n_pilots, const double* M, double removed_max);
int chest_set_estimator_factors(uint64_t handle, int scheme, int variant, int n_snr, const double* R_inv, double removed_max);
int chest_estimator_info(uint64_t handle, int scheme, int* mode, int* factored, double* removed_r, double* removed_w, float* ms);
/* Channel-estimation error sums next to the bit-error counters (north_star: "BER/MSE counters"; the reference itself keeps
 * no MSE): with accumulation enabled every chest_run_batch* also leaves sum_i |h_est(i) - h(i)|^2 over the K positions of
 * a scheme's grid -- h_est = diag(D_est) of DS.m:428/517, h = diag(Q'HG) of DS.m:392-393 -- for every realization, SNR
 * point and iteration (0 = the one-tap stage).  chest_get_mse: out[rep][snr][it][scheme], n_rep x n_snr x (n_iter+1) x 3
 * doubles of the last batch; divide by K for the mean.  Off by default (one extra load per estimated value). */
int chest_set_mse_accumulation(uint64_t handle, int enable);
int chest_get_mse(uint64_t handle, double* out);
/* Work units (up to 16 columns x all K rows each) the IC kernels processed in the last batch. */
int chest_unit_count(uint64_t handle, int* n_units);

/* Same with the result left on the device: err_dev is a device pointer (used by bench `value`). */
int chest_run_batch_device(uint64_t handle, int n_rep, int n_iter, const chest_draws* draws,
                           uint64_t seed, int64_t first_rep, uint32_t* err_dev);

/* ------------------------------------------------------------------ one host thread, several GPUs
 * (SURVEY.md 8b "Threading", 8e: a MEX host is single-threaded; realizations are independent, DS.m:350-368)
 *
 * chest_run_batch_async enqueues the loop body for n_rep realizations on the context's stream and returns without
 * waiting; chest_wait blocks until it has finished and copies the error counts (layout of chest_run_batch) from a
 * pinned buffer the context owns into err_out (may be NULL).  A caller with one context per device enqueues on all
 * of them, then waits for all of them.  Seeded draws (draws == NULL) or device-resident draws never block the host;
 * explicit host draws are staged by the CUDA runtime unless they live in pinned memory. */
int chest_run_batch_async(uint64_t handle, int n_rep, int n_iter, const chest_draws* draws, uint64_t seed,
                          int64_t first_rep);
int chest_wait(uint64_t handle, uint32_t* err_out);

/* A group of finalized contexts, one per device, configured identically (the caller runs the same setters on each).
 * chest_multi_run shards n_rep_total seeded realizations [first_rep, first_rep + n_rep_total) over the devices in
 * contiguous blocks (rounds of at most max_batch per device, all devices of a round enqueued before any wait; no
 * inter-GPU traffic during the loop), sums each device's counters on that device (64-bit), and finishes with the one
 * collective of the path: an NCCL all-reduce (ncclUint64, sum; NCCL bound with dlopen, ncclCommInitAll) of the
 * S x (n_iter+1) x 12 totals.  err_out (may be NULL): per-realization counts [n_rep_total][snr][it][12];
 * totals_out (may be NULL): [snr][it][scheme][csi][edge] uint64; reduce_ms (may be NULL): device time of the
 * all-reduce, maximum over the devices.  The contexts stay owned by the caller. */
int chest_multi_create(const uint64_t* handles, int n_devices, uint64_t* multi);
int chest_multi_run(uint64_t multi, int64_t n_rep_total, int n_iter, uint64_t seed, int64_t first_rep,
                    uint32_t* err_out, uint64_t* totals_out, float* reduce_ms);
int chest_multi_destroy(uint64_t multi);

/* n_bits[scheme][edge] used as BER denominators. */
int chest_bit_counts(uint64_t handle, int64_t* n_bits /* [CHEST_N_SCHEMES][2] */);

/* Fill device-resident draws for n_rep realizations with the counter-based generator; the
 * returned struct points into context-owned device memory (valid until the next call). */
int chest_generate_draws(uint64_t handle, int n_rep, uint64_t seed, int64_t first_rep, chest_draws* out);
/* Copy generated draws back to host buffers laid out like chest_draws (any pointer may be NULL). */
int chest_download_draws(uint64_t handle, int n_rep, double* doppler_u, double* phase_u,
                         uint8_t* bits_aux, uint8_t* bits_cod, uint8_t* bits_ofdm,
                         int32_t* pilot_idx_fbmc, int32_t* pilot_idx_ofdm, double* noise);

/* Parity-mode read-back of per-realization state left by the last chest_run_batch:
 * what = 0: y (K), 1: hP of the last iteration (P), 2: data-symbol estimates of the last iteration,
 * estimated CSI (n_data), 3: same, perfect CSI (n_data), 4: h_est = diag(D_est) of the last
 * iteration (K).  Complex interleaved. */
int chest_get_state(uint64_t handle, int what, int scheme, int rep, int i_snr, double* out);

/* Kernel-launch counter (all kernels launched by this context since creation). */
int64_t chest_launch_count(uint64_t handle);
/* Per-stage device time of the last chest_run_batch in ms (CUDA events on the context stream):
 * [0] draws/rng, [1] K1 channel+TX, [2] K2 transmission matrix, [3] demodulation,
 * [4] one-tap stage, [5] IC iterations, [6] total.  Requires chest_set_profiling(handle, 1). */
int chest_set_profiling(uint64_t handle, int enable);
int chest_stage_times(uint64_t handle, float* ms /* [7] */);
/* K1 banded operator applied to G (k_apply_hg) in the last profiled chest_run_batch: device time (ms) and
 * algorithmic bytes (H*G rows written + h read), for the HBM roofline of the banded path. */
int chest_banded_apply_stats(uint64_t handle, float* ms, double* bytes);
/* Device time (ms, CUDA events on the context's stream) of the hot kernels in the last profiled chest_run_batch:
 * [0] k_apply_hg, [1] k_gemm_d (K2; both zero in factored mode), [2] k_ic_main summed over the iterations,
 * [3] k_ic_light summed, [4] the factored perfect-CSI chain summed, [5] the diag(D) GEMM, [6] k_synth_h (channel
 * synthesis), [7] k_tx_symbols, [8] s = G x (both waveforms), [9] k_apply_h (r0 = H s, both waveforms). */
int chest_kernel_times(uint64_t handle, float* ms /* [10] */);
/* Algorithmic work of one realization for the roofline:
 * [0] K2 support-aware flops, [1] K3/K4 estimated-CSI flops per iteration-evaluation set,
 * [2] perfect-CSI flops, [3] demod/TX flops, [4] bytes of W streamed per IC kernel launch,
 * [5] precoding flops, [6] the part of [1] executed by k_ic_main (off-diagonal products),
 * [7] flops of the factored perfect-CSI chain (G v, H, Q^H over the supports, all iterations). */
int chest_work_model(uint64_t handle, int n_iter, double* out /* [8] */);

/* Device-timeline timing for callers that cannot see the context's stream: record event `slot`
 * (0..3) on it; elapsed time between two recorded slots in ms (synchronises on the later one). */
int chest_event_record(uint64_t handle, int slot);
int chest_event_elapsed(uint64_t handle, int slot_a, int slot_b, float* ms);

/* FP64 peak probe: runs a register-resident DMMA (mode 0) or DFMA (mode 1) loop on every SM for
 * `iters` iterations and returns achieved TFLOP/s; the roofline denominator for K2-K4. */
int chest_fp64_peak(uint64_t handle, int mode, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CHEST_B200_H */
