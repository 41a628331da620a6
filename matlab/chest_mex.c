/* chest_mex.c -- MATLAB / Octave MEX gateway over the C ABI of include/chest_b200.h.
 *
 *   out = chest_mex('command', handle, args...)
 *
 * One gateway, dispatch on a command string; the uint64 context handle lives in the MATLAB
 * classdef wrappers (matlab/+Channel/FastFading.m, matlab/+Modulation/FBMC.m, OFDM.m, matlab/+ChestB200/Simulation.m);
 * every live context is also remembered here and released by the mexAtExit hook.  The shim only
 * converts between MATLAB's split-complex, 1-based, column-major arrays and the ABI's interleaved,
 * 0-based ones; launchers return codes and errors are raised AFTER temporaries are released
 * (mexErrMsgIdAndTxt does not return).  No gpuArray, no Parallel Computing Toolbox, no CPU path.
 * Build (with MATLAB):  mex -I../include chest_mex.c -L../channel-estimation_b200 -lchest_b200
 * In this repository it is compile-checked against matlab/stub/mex.h only (no MATLAB in the image). */
#include <string.h>
#include "mex.h"
#include "chest_b200.h"

/* Every live context is remembered so that 'clear mex' / MATLAB exit releases the device memory (mexAtExit). */
#define CHEST_MEX_MAX_HANDLES 64
static uint64_t g_handles[CHEST_MEX_MAX_HANDLES];
static uint64_t g_multi[CHEST_MEX_MAX_HANDLES];
static int g_at_exit_registered = 0;
static void at_exit_cleanup(void) {
    int i;
    for (i = 0; i < CHEST_MEX_MAX_HANDLES; ++i) if (g_multi[i]) { chest_multi_destroy(g_multi[i]); g_multi[i] = 0; }
    for (i = 0; i < CHEST_MEX_MAX_HANDLES; ++i) if (g_handles[i]) { chest_destroy(g_handles[i]); g_handles[i] = 0; }
}
static void remember(uint64_t* table, uint64_t h) {
    int i;
    if (!g_at_exit_registered) { mexAtExit(at_exit_cleanup); g_at_exit_registered = 1; }
    for (i = 0; i < CHEST_MEX_MAX_HANDLES; ++i) if (!table[i]) { table[i] = h; return; }
}
static void forget(uint64_t* table, uint64_t h) {
    int i;
    for (i = 0; i < CHEST_MEX_MAX_HANDLES; ++i) if (table[i] == h) table[i] = 0;
}

static void fail_if(int rc) {
    if (rc != CHEST_OK) mexErrMsgIdAndTxt("chest:error", "%s", chest_last_error());
}
static uint64_t handle_of(const mxArray* a) { return *(const uint64_t*)mxGetData(a); }

/* MATLAB split complex -> interleaved (re,im) pairs */
static double* interleave(const mxArray* a) {
    size_t n = mxGetNumberOfElements(a), i;
    double* out = (double*)mxMalloc(2 * n * sizeof(double) + 16);
    const double *re = mxGetPr(a), *im = mxIsComplex(a) ? mxGetPi(a) : NULL;
    for (i = 0; i < n; ++i) { out[2 * i] = re[i]; out[2 * i + 1] = im ? im[i] : 0.0; }
    return out;
}
static mxArray* deinterleave(const double* z, mwSize m, mwSize n) {
    mxArray* a = mxCreateDoubleMatrix(m, n, mxCOMPLEX);
    double *re = mxGetPr(a), *im = mxGetPi(a);
    size_t i, tot = (size_t)m * n;
    for (i = 0; i < tot; ++i) { re[i] = z[2 * i]; im[i] = z[2 * i + 1]; }
    return a;
}
static int32_t* to_i32_zero_based(const mxArray* a) {      /* 1-based doubles -> 0-based int32 */
    size_t n = mxGetNumberOfElements(a), i;
    int32_t* out = (int32_t*)mxMalloc(n * sizeof(int32_t) + 4);
    const double* p = mxGetPr(a);
    for (i = 0; i < n; ++i) out[i] = (int32_t)p[i] - 1;
    return out;
}
static uint8_t* to_u8(const mxArray* a) {
    size_t n = mxGetNumberOfElements(a), i;
    uint8_t* out = (uint8_t*)mxMalloc(n + 1);
    const double* p = mxGetPr(a);
    for (i = 0; i < n; ++i) out[i] = p[i] != 0.0;
    return out;
}
static int32_t* to_i32(const mxArray* a, int offset) {      /* doubles -> int32 (offset -1 for 1-based indices) */
    size_t n = mxGetNumberOfElements(a), i;
    int32_t* out = (int32_t*)mxMalloc(n * sizeof(int32_t) + 4);
    const double* p = mxGetPr(a);
    for (i = 0; i < n; ++i) out[i] = (int32_t)p[i] + offset;
    return out;
}
static mxArray* handle_array(uint64_t h) {
    mxArray* a = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t*)mxGetData(a) = h;
    return a;
}
/* sparse complex matrix -> CSC arrays of the ABI */
static void sparse_parts(const mxArray* a, int64_t** jc, int64_t** ir, double** val) {
    mwSize ncol = mxGetN(a), c;
    const mwIndex *Jc = mxGetJc(a), *Ir = mxGetIr(a);
    mwIndex nnz = Jc[ncol], e;
    const double *re = mxGetPr(a), *im = mxIsComplex(a) ? mxGetPi(a) : NULL;
    *jc = (int64_t*)mxMalloc((ncol + 1) * sizeof(int64_t));
    *ir = (int64_t*)mxMalloc((nnz + 1) * sizeof(int64_t));
    *val = (double*)mxMalloc((2 * nnz + 2) * sizeof(double));
    for (c = 0; c <= ncol; ++c) (*jc)[c] = (int64_t)Jc[c];
    for (e = 0; e < nnz; ++e) { (*ir)[e] = (int64_t)Ir[e]; (*val)[2 * e] = re[e]; (*val)[2 * e + 1] = im ? im[e] : 0.0; }
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[64];
    int rc = CHEST_OK;
    (void)nlhs;
    if (nrhs < 1 || !mxIsChar(prhs[0]) || mxGetString(prhs[0], cmd, sizeof(cmd)))
        mexErrMsgIdAndTxt("chest:usage", "chest_mex('command', ...)");

    if (!strcmp(cmd, "create")) {                        /* h = chest_mex('create', device) */
        uint64_t h = 0;
        rc = chest_create(nrhs > 1 ? (int)mxGetScalar(prhs[1]) : 0, &h);
        fail_if(rc);
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t*)mxGetData(plhs[0]) = h;
        remember(g_handles, h);
        mexLock();                                       /* device state must survive 'clear mex' */
    } else if (!strcmp(cmd, "destroy")) {
        rc = chest_destroy(handle_of(prhs[1]));
        forget(g_handles, handle_of(prhs[1]));
        mexUnlock();
        fail_if(rc);
    } else if (!strcmp(cmd, "set_channel")) {            /* (h, N, pdp_normalized, fD, dt, paths, model) */
        rc = chest_set_channel(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetNumberOfElements(prhs[3]),
                               mxGetPr(prhs[3]), mxGetScalar(prhs[4]), mxGetScalar(prhs[5]), (int)mxGetScalar(prhs[6]),
                               (int)mxGetScalar(prhs[7]));
        fail_if(rc);
    } else if (!strcmp(cmd, "set_waveform")) {           /* (h, wf, G, Q) : G = GetTXMatrix, Q = GetRXMatrix' */
        double *G = interleave(prhs[3]), *Q = interleave(prhs[4]);
        rc = chest_set_waveform(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetM(prhs[3]), (int)mxGetN(prhs[3]), G, Q);
        mxFree(G); mxFree(Q);
        fail_if(rc);
    } else if (!strcmp(cmd, "set_constellation")) {      /* (h, which, SymbolMapping, BitMapping) */
        double* s = interleave(prhs[3]);
        uint8_t* b = to_u8(prhs[4]);
        rc = chest_set_constellation(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetNumberOfElements(prhs[3]), s, b);
        mxFree(s); mxFree(b);
        fail_if(rc);
    } else if (!strcmp(cmd, "set_scheme")) {
        /* (h, scheme, wf, sparse(C), pilot_pos, data_pos|[], kappa, dpr, detect_mode, constellation, ConsideredBits) */
        int64_t *jc, *ir64; double* val; int32_t *ir, *pp, *dp = NULL; uint8_t* cb; mwIndex e, nnz;
        sparse_parts(prhs[4], &jc, &ir64, &val);
        nnz = (mwIndex)jc[mxGetN(prhs[4])];
        ir = (int32_t*)mxMalloc((nnz + 1) * sizeof(int32_t));
        for (e = 0; e < nnz; ++e) ir[e] = (int32_t)ir64[e];
        pp = to_i32_zero_based(prhs[5]);
        if (!mxIsEmpty(prhs[6])) dp = to_i32_zero_based(prhs[6]);
        cb = to_u8(prhs[11]);
        rc = chest_set_scheme(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), (int)mxGetN(prhs[4]),
                              (int)mxGetNumberOfElements(prhs[5]), (int)(mxGetN(prhs[4]) - mxGetNumberOfElements(prhs[5])),
                              jc, ir, val, pp, dp, mxGetScalar(prhs[7]), mxGetScalar(prhs[8]), (int)mxGetScalar(prhs[9]),
                              (int)mxGetScalar(prhs[10]), cb);
        mxFree(jc); mxFree(ir64); mxFree(val); mxFree(ir); mxFree(pp); if (dp) mxFree(dp); mxFree(cb);
        fail_if(rc);
    } else if (!strcmp(cmd, "set_snr")) {                /* (h, Pn_time vector) */
        rc = chest_set_snr(handle_of(prhs[1]), (int)mxGetNumberOfElements(prhs[2]), mxGetPr(prhs[2]));
        fail_if(rc);
    } else if (!strcmp(cmd, "set_mmse")) {               /* (h, scheme, variant, W_MMSE sparse K^2P x S) */
        int64_t *jc, *ir; double* val;
        sparse_parts(prhs[4], &jc, &ir, &val);
        rc = chest_set_mmse(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), (int)mxGetN(prhs[4]), jc, ir, val);
        mxFree(jc); mxFree(ir); mxFree(val);
        fail_if(rc);
    } else if (!strcmp(cmd, "finalize")) {               /* (h, max_batch) */
        rc = chest_finalize(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]));
        fail_if(rc);
    } else if (!strcmp(cmd, "new_realization")) {        /* (h, batch, seed, first_rep) */
        rc = chest_new_realization_seeded(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (uint64_t)mxGetScalar(prhs[3]),
                                          (int64_t)mxGetScalar(prhs[4]));
        fail_if(rc);
    } else if (!strcmp(cmd, "impulse_response")) {       /* h_out = (h, b, N, Lt) */
        mwSize N = (mwSize)mxGetScalar(prhs[3]), Lt = (mwSize)mxGetScalar(prhs[4]);
        double* z = (double*)mxMalloc(2 * N * Lt * sizeof(double));
        rc = chest_get_impulse_response(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), z);
        if (rc == CHEST_OK) plhs[0] = deinterleave(z, N, Lt);
        mxFree(z);
        fail_if(rc);
    } else if (!strcmp(cmd, "convolution_matrix")) {     /* cell = (h, b, N): {1,1} sparse N x N */
        mwSize N = (mwSize)mxGetScalar(prhs[3]);
        int64_t nnz = 0, *jc; int32_t* ir; double* val; mxArray* S; mwIndex e;
        rc = chest_get_convolution_csc(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), &nnz, NULL, NULL, NULL);
        fail_if(rc);
        jc = (int64_t*)mxMalloc((N + 1) * sizeof(int64_t)); ir = (int32_t*)mxMalloc((nnz + 1) * sizeof(int32_t));
        val = (double*)mxMalloc((2 * nnz + 2) * sizeof(double));
        rc = chest_get_convolution_csc(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), &nnz, jc, ir, val);
        if (rc == CHEST_OK) {
            S = mxCreateSparse(N, N, (mwSize)nnz, mxCOMPLEX);
            for (e = 0; e <= N; ++e) mxGetJc(S)[e] = (mwIndex)jc[e];
            for (e = 0; e < (mwIndex)nnz; ++e) { mxGetIr(S)[e] = (mwIndex)ir[e]; mxGetPr(S)[e] = val[2 * e]; mxGetPi(S)[e] = val[2 * e + 1]; }
            plhs[0] = mxCreateCellMatrix(1, 1);
            mxSetCell(plhs[0], 0, S);
        }
        mxFree(jc); mxFree(ir); mxFree(val);
        fail_if(rc);
    } else if (!strcmp(cmd, "convolve") || !strcmp(cmd, "modulate") || !strcmp(cmd, "demodulate")) {
        /* r = (h, b|wf, s, n_out): column vectors in, column vectors out */
        double* in = interleave(prhs[3]);
        mwSize ncol = mxGetN(prhs[3]), nout = (mwSize)mxGetScalar(prhs[4]);
        double* out = (double*)mxMalloc(2 * nout * ncol * sizeof(double));
        if (!strcmp(cmd, "convolve")) rc = chest_convolve(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), in, (int)ncol, out);
        else if (!strcmp(cmd, "modulate")) rc = chest_modulate(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), in, (int)ncol, out);
        else rc = chest_demodulate(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), in, (int)ncol, out);
        if (rc == CHEST_OK) plhs[0] = deinterleave(out, nout, ncol);
        mxFree(in); mxFree(out);
        fail_if(rc);
    } else if (!strcmp(cmd, "transmission_matrix")) {    /* [D, hdiag] = (h, b, wf, K) */
        mwSize K = (mwSize)mxGetScalar(prhs[4]);
        double *D = (double*)mxMalloc(2 * K * K * sizeof(double)), *hd = (double*)mxMalloc(2 * K * sizeof(double));
        rc = chest_transmission_matrix(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), D, hd);
        if (rc == CHEST_OK) { plhs[0] = deinterleave(D, K, K); if (nlhs > 1) plhs[1] = deinterleave(hd, K, 1); }
        mxFree(D); mxFree(hd);
        fail_if(rc);
    } else if (!strcmp(cmd, "estimate")) {               /* [Dhat, hdiag] = (h, scheme, variant, i_snr, hP, K) */
        mwSize K = (mwSize)mxGetScalar(prhs[6]);
        double* hP = interleave(prhs[5]);
        double *D = (double*)mxMalloc(2 * K * K * sizeof(double)), *hd = (double*)mxMalloc(2 * K * sizeof(double));
        rc = chest_estimate(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), (int)mxGetScalar(prhs[4]) - 1, hP, D, hd);
        if (rc == CHEST_OK) { plhs[0] = deinterleave(D, K, K); if (nlhs > 1) plhs[1] = deinterleave(hd, K, 1); }
        mxFree(hP); mxFree(D); mxFree(hd);
        fail_if(rc);
    } else if (!strcmp(cmd, "run_batch")) {              /* err = (h, n_rep, n_iter, seed, first_rep, n_snr) */
        mwSize n_rep = (mwSize)mxGetScalar(prhs[2]), n_iter = (mwSize)mxGetScalar(prhs[3]), n_snr = (mwSize)mxGetScalar(prhs[6]);
        plhs[0] = mxCreateNumericMatrix(12 * (n_iter + 1) * n_snr, n_rep, mxUINT32_CLASS, mxREAL);
        rc = chest_run_batch(handle_of(prhs[1]), (int)n_rep, (int)n_iter, NULL, (uint64_t)mxGetScalar(prhs[4]),
                             (int64_t)mxGetScalar(prhs[5]), (uint32_t*)mxGetData(plhs[0]));
        fail_if(rc);
    } else if (!strcmp(cmd, "device_info")) {            /* [n_sm, cc] = (device) */
        int n_sm = 0, maj = 0, min_ = 0;
        rc = chest_device_info(nrhs > 1 ? (int)mxGetScalar(prhs[1]) : 0, &n_sm, &maj, &min_);
        fail_if(rc);
        plhs[0] = mxCreateDoubleScalar((double)n_sm);
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(maj * 10.0 + min_);
    } else if (!strcmp(cmd, "channel_info")) {           /* [n_shift, fD, T] = (h) */
        int ns = 0, nt = 0; double fd = 0.0;
        rc = chest_channel_info(handle_of(prhs[1]), &ns, &fd, &nt);
        fail_if(rc);
        plhs[0] = mxCreateDoubleScalar((double)ns);
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(fd);
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar((double)nt);
    } else if (!strcmp(cmd, "new_realization_draws")) {  /* (h, batch, doppler_u, phase_u): rand([T 1 Paths]) order per column */
        rc = chest_new_realization(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]));
        fail_if(rc);
    } else if (!strcmp(cmd, "new_realization_gauss")) {  /* (h, batch, gauss): 'Discrete-*' normals, (2 n_shift + 1) x T per column */
        double* g = interleave(prhs[3]);
        rc = chest_new_realization_gauss(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), g);
        mxFree(g);
        fail_if(rc);
    } else if (!strcmp(cmd, "set_impulse_response")) {   /* (h, batch, h_in): N x Lt x batch complex */
        double* z = interleave(prhs[3]);
        rc = chest_set_impulse_response(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), z);
        mxFree(z);
        fail_if(rc);
    } else if (!strcmp(cmd, "set_modem")) {
        /* (h, wf, kind, L, K, FFTSize, bins(0-based), TimeSpacing, O, CP, ZeroGuard, PrototypeFilter|[], PhaseShift|[], NormalizationFactor, F) */
        int32_t* bins = to_i32(prhs[7], 0);
        double* ph = mxIsEmpty(prhs[13]) ? NULL : interleave(prhs[13]);
        rc = chest_set_modem(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), (int)mxGetScalar(prhs[4]),
                             (int)mxGetScalar(prhs[5]), (int)mxGetScalar(prhs[6]), bins, (int)mxGetScalar(prhs[8]),
                             (int)mxGetScalar(prhs[9]), (int)mxGetScalar(prhs[10]), (int)mxGetScalar(prhs[11]),
                             mxIsEmpty(prhs[12]) ? NULL : mxGetPr(prhs[12]), ph, mxGetScalar(prhs[14]), mxGetScalar(prhs[15]));
        mxFree(bins); if (ph) mxFree(ph);
        fail_if(rc);
    } else if (!strcmp(cmd, "modulate_fft") || !strcmp(cmd, "demodulate_fft")) {   /* out = (h, wf, in, n_out_rows) */
        double* in = interleave(prhs[3]);
        mwSize ncol = mxGetN(prhs[3]), nout = (mwSize)mxGetScalar(prhs[4]);
        double* out = (double*)mxMalloc(2 * nout * ncol * sizeof(double));
        if (!strcmp(cmd, "modulate_fft")) rc = chest_modulate_fft(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), in, (int)ncol, out);
        else rc = chest_demodulate_fft(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), in, (int)ncol, out);
        if (rc == CHEST_OK) plhs[0] = deinterleave(out, nout, ncol);
        mxFree(in); mxFree(out);
        fail_if(rc);
    } else if (!strcmp(cmd, "setup_correlations")) {     /* [R_hP, n_support] = (h, wf, pilot_pos(1-based), TimeCorrelation, threshold) */
        mwSize P = mxGetNumberOfElements(prhs[3]);
        int32_t* pp = to_i32(prhs[3], -1);
        double* R = (double*)mxMalloc(2 * P * P * sizeof(double));
        int64_t nsup = 0;
        rc = chest_setup_correlations(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)P, pp, mxGetPr(prhs[4]), mxGetScalar(prhs[5]), R, &nsup);
        if (rc == CHEST_OK) { plhs[0] = deinterleave(R, P, P); if (nlhs > 1) plhs[1] = mxCreateDoubleScalar((double)nsup); }
        mxFree(pp); mxFree(R);
        fail_if(rc);
    } else if (!strcmp(cmd, "build_mmse")) {             /* (h, scheme, variant, R_inv: P x P x n_snr complex, threshold) */
        mwSize P = mxGetM(prhs[4]);
        double* Ri = interleave(prhs[4]);
        rc = chest_build_mmse(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]),
                              (int)(mxGetNumberOfElements(prhs[4]) / (P * P)), Ri, mxGetScalar(prhs[5]));
        mxFree(Ri);
        fail_if(rc);
    } else if (!strcmp(cmd, "release_setup")) {
        rc = chest_release_setup(handle_of(prhs[1]));
        fail_if(rc);
    } else if (!strcmp(cmd, "set_perfect_csi_mode")) {   /* (h, mode): 0 dense D, 1 factored */
        rc = chest_set_perfect_csi_mode(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]));
        fail_if(rc);
    } else if (!strcmp(cmd, "set_precision")) {          /* (h, mode): 0 FP64, 1 split-BF16 tensor-core mode (stated 1e-4) */
        rc = chest_set_precision(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]));
        fail_if(rc);
    } else if (!strcmp(cmd, "set_estimator_mode")) {     /* (h, mode): 0 auto, 1 thresholded W tiles, 2 factored (stated tolerance), 3 factored where exact */
        rc = chest_set_estimator_mode(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]));
        fail_if(rc);
    } else if (!strcmp(cmd, "estimator_info")) {         /* [factored, removed_R, removed_W] = (h, scheme) */
        int factored = 0; double rr = 0, rw = 0;
        rc = chest_estimator_info(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), NULL, &factored, &rr, &rw, NULL);
        fail_if(rc);
        plhs[0] = mxCreateDoubleScalar((double)factored);
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(rr);
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(rw);
    } else if (!strcmp(cmd, "run_batch_draws")) {
        /* err = (h, n_rep, n_iter, n_snr, doppler_u, phase_u, bits_aux, bits_cod, bits_ofdm, pilot_idx_fbmc, pilot_idx_ofdm, noise)
         * one column per realization, in the order DS.m:352-368,399 draws them; pilot indices 1-based; noise N x n_snr x n_rep */
        mwSize n_rep = (mwSize)mxGetScalar(prhs[2]), n_iter = (mwSize)mxGetScalar(prhs[3]), n_snr = (mwSize)mxGetScalar(prhs[4]);
        chest_draws d;
        uint8_t* b[3] = {NULL, NULL, NULL}; int32_t* pi_[2] = {NULL, NULL}; double* nz; int k;
        memset(&d, 0, sizeof(d));
        d.doppler_u = mxGetPr(prhs[5]); d.phase_u = mxGetPr(prhs[6]);
        for (k = 0; k < 3; ++k) if (!mxIsEmpty(prhs[7 + k])) { b[k] = to_u8(prhs[7 + k]); d.bits[k] = b[k]; }
        for (k = 0; k < 2; ++k) if (!mxIsEmpty(prhs[10 + k])) { pi_[k] = to_i32(prhs[10 + k], -1); d.pilot_idx[k] = pi_[k]; }
        nz = interleave(prhs[12]); d.noise = nz;
        plhs[0] = mxCreateNumericMatrix(12 * (n_iter + 1) * n_snr, n_rep, mxUINT32_CLASS, mxREAL);
        rc = chest_run_batch(handle_of(prhs[1]), (int)n_rep, (int)n_iter, &d, 0, 0, (uint32_t*)mxGetData(plhs[0]));
        for (k = 0; k < 3; ++k) if (b[k]) mxFree(b[k]);
        for (k = 0; k < 2; ++k) if (pi_[k]) mxFree(pi_[k]);
        mxFree(nz);
        fail_if(rc);
    } else if (!strcmp(cmd, "run_batch_async")) {        /* (h, n_rep, n_iter, seed, first_rep): returns at once */
        rc = chest_run_batch_async(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), NULL,
                                   (uint64_t)mxGetScalar(prhs[4]), (int64_t)mxGetScalar(prhs[5]));
        fail_if(rc);
    } else if (!strcmp(cmd, "wait")) {                   /* err = (h, n_rep, n_iter, n_snr) */
        mwSize n_rep = (mwSize)mxGetScalar(prhs[2]), n_iter = (mwSize)mxGetScalar(prhs[3]), n_snr = (mwSize)mxGetScalar(prhs[4]);
        plhs[0] = mxCreateNumericMatrix(12 * (n_iter + 1) * n_snr, n_rep, mxUINT32_CLASS, mxREAL);
        rc = chest_wait(handle_of(prhs[1]), (uint32_t*)mxGetData(plhs[0]));
        fail_if(rc);
    } else if (!strcmp(cmd, "multi_create")) {           /* m = (handles: uint64 vector, one finalized context per device) */
        uint64_t m = 0;
        rc = chest_multi_create((const uint64_t*)mxGetData(prhs[1]), (int)mxGetNumberOfElements(prhs[1]), &m);
        fail_if(rc);
        plhs[0] = handle_array(m);
        remember(g_multi, m);
    } else if (!strcmp(cmd, "multi_run")) {              /* [err, totals, reduce_ms] = (m, n_rep_total, n_iter, seed, first_rep, n_snr) */
        mwSize n_rep = (mwSize)mxGetScalar(prhs[2]), n_iter = (mwSize)mxGetScalar(prhs[3]), n_snr = (mwSize)mxGetScalar(prhs[6]);
        float ms = 0.0f;
        mxArray* tot = mxCreateNumericMatrix(12 * (n_iter + 1) * n_snr, 1, mxUINT64_CLASS, mxREAL);
        plhs[0] = mxCreateNumericMatrix(12 * (n_iter + 1) * n_snr, n_rep, mxUINT32_CLASS, mxREAL);
        rc = chest_multi_run(handle_of(prhs[1]), (int64_t)n_rep, (int)n_iter, (uint64_t)mxGetScalar(prhs[4]),
                             (int64_t)mxGetScalar(prhs[5]), (uint32_t*)mxGetData(plhs[0]), (uint64_t*)mxGetData(tot), &ms);
        fail_if(rc);
        if (nlhs > 1) plhs[1] = tot;
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar((double)ms);
    } else if (!strcmp(cmd, "multi_destroy")) {
        rc = chest_multi_destroy(handle_of(prhs[1]));
        forget(g_multi, handle_of(prhs[1]));
        fail_if(rc);
    } else if (!strcmp(cmd, "set_interpolation")) {      /* (h, scheme, InterpolationMatrix K x P) */
        double* M = interleave(prhs[3]);
        rc = chest_set_interpolation(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), M);
        mxFree(M);
        fail_if(rc);
    } else if (!strcmp(cmd, "sv_run_batch")) {           /* err = (h, n_body, Pn_time vector (one per body), seed, first_body): 5 x n_body */
        mwSize n_body = (mwSize)mxGetScalar(prhs[2]);
        plhs[0] = mxCreateNumericMatrix(5, n_body, mxUINT32_CLASS, mxREAL);
        rc = chest_sv_run_batch(handle_of(prhs[1]), (int)n_body, mxGetPr(prhs[3]), NULL, (uint64_t)mxGetScalar(prhs[4]),
                                (int64_t)mxGetScalar(prhs[5]), (uint32_t*)mxGetData(plhs[0]));
        fail_if(rc);
    } else if (!strcmp(cmd, "bit_counts")) {             /* n = (h): 2 x 3 (edge x scheme) */
        int64_t nb[6]; int i;
        rc = chest_bit_counts(handle_of(prhs[1]), nb);
        fail_if(rc);
        plhs[0] = mxCreateDoubleMatrix(2, 3, mxREAL);
        for (i = 0; i < 6; ++i) mxGetPr(plhs[0])[i] = (double)nb[i];
    } else {
        mexErrMsgIdAndTxt("chest:usage", "unknown command '%s'", cmd);
    }
}
