/* chest_mex.c -- MATLAB / Octave MEX gateway over the C ABI of include/chest_b200.h.
 *
 *   out = chest_mex('command', handle, args...)
 *
 * One gateway, dispatch on a command string; the uint64 context handle lives in the MATLAB
 * classdef wrappers (matlab/+Channel/FastFading.m, matlab/+ChestB200/Simulation.m).  The shim only
 * converts between MATLAB's split-complex, 1-based, column-major arrays and the ABI's interleaved,
 * 0-based ones; launchers return codes and errors are raised AFTER temporaries are released
 * (mexErrMsgIdAndTxt does not return).  No gpuArray, no Parallel Computing Toolbox, no CPU path.
 * Build (with MATLAB):  mex -I../include chest_mex.c -L../channel-estimation_b200 -lchest_b200
 * In this repository it is compile-checked against matlab/stub/mex.h only (no MATLAB in the image). */
#include <string.h>
#include "mex.h"
#include "chest_b200.h"

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
        mexLock();                                       /* device state must survive 'clear mex' */
    } else if (!strcmp(cmd, "destroy")) {
        rc = chest_destroy(handle_of(prhs[1]));
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
