"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical
inputs.  Tolerances: FP64 quantities 1e-9 relative (north_star), hard decisions / bit-error
counts identical."""
import numpy as np
import pytest

from tests.helpers import err_from_oracle

pytestmark = pytest.mark.gpu

TOL = 1e-9


def rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.fixture(scope="module")
def draws3(ds_default):
    from oracle.ds import new_draws
    rng = np.random.default_rng(2024)
    return [new_draws(ds_default, rng) for _ in range(3)]


def test_k1_impulse_response_and_convolution(ds_default, gpu_ctx, draws3):
    S, ctx = ds_default, gpu_ctx
    du = np.stack([d["doppler_u"].reshape(-1, order="F") for d in draws3])
    pu = np.stack([d["phase_u"].reshape(-1, order="F") for d in draws3])
    ctx.new_realization(du, pu)
    chan = S["chan"]
    rng = np.random.default_rng(5)
    for b, d in enumerate(draws3):
        chan.NewRealization(d["doppler_u"], d["phase_u"])
        h = ctx.impulse_response(b)
        assert h.shape == chan.ImpulseResponse.shape
        assert rel(h, chan.ImpulseResponse) < 1e-12
        H_ref = chan.GetConvolutionMatrix()
        H = ctx.convolution_matrix(b)
        assert H.nnz == H_ref.nnz == sum(S["N"] - m for m in chan.Implementation["IndexDelayTaps"])
        assert np.array_equal(H.indptr, H_ref.indptr) and np.array_equal(H.indices, H_ref.indices)
        assert rel(H.data, H_ref.data) < 1e-12
        s = rng.standard_normal((S["N"], 2)) + 1j * rng.standard_normal((S["N"], 2))
        assert rel(ctx.convolve(s, b), H_ref @ s) < 1e-12
        assert rel(ctx.convolve(s[:, 0], b), chan.Convolution(s[:, 0])) < 1e-12


def test_k2_transmission_matrix(ds_default, gpu_ctx, draws3):
    S, ctx = ds_default, gpu_ctx
    d = draws3[1]
    ctx.new_realization(d["doppler_u"].reshape(1, -1, order="F"), d["phase_u"].reshape(1, -1, order="F"))
    S["chan"].NewRealization(d["doppler_u"], d["phase_u"])
    H = S["chan"].GetConvolutionMatrix()
    for wf in ("F", "O"):
        w = S["wf"][wf]
        D_ref = w["Q"].conj().T @ (H @ w["G"])
        D, h = ctx.transmission_matrix(wf, 0)
        assert rel(D, D_ref) < TOL
        assert rel(h, np.diag(D_ref)) < TOL
        assert np.array_equal(h, np.diag(D))


def test_k2_uploaded_impulse_response(ds_default, gpu_ctx):
    """Exported channel realizations (or pseudo-channels) go through the same K2."""
    S, ctx = ds_default, gpu_ctx
    rng = np.random.default_rng(11)
    Lt = len(S["chan"].Implementation["PowerDelayProfileNormalized"])
    h = rng.standard_normal((2, S["N"], Lt)) + 1j * rng.standard_normal((2, S["N"], Lt))
    ctx.set_impulse_response(h)
    import scipy.sparse as sp
    for b in range(2):
        rows = np.concatenate([np.arange(m, S["N"]) for m in range(Lt)])
        cols = np.concatenate([np.arange(m, S["N"]) - m for m in range(Lt)])
        vals = np.concatenate([h[b, m:, m] for m in range(Lt)])
        H = sp.csc_matrix((vals, (rows, cols)), shape=(S["N"], S["N"]))
        w = S["wf"]["O"]
        D, _ = ctx.transmission_matrix("O", b)
        assert rel(D, w["Q"].conj().T @ (H @ w["G"])) < TOL


def test_modem_matrix_form(ds_default, gpu_ctx):
    S, ctx = ds_default, gpu_ctx
    rng = np.random.default_rng(3)
    for wf, modem in (("F", S["fbmc"]), ("O", S["ofdm"])):
        w = S["wf"][wf]
        L, Ksym = modem.Nr["Subcarriers"], modem.Nr["MCSymbols"]
        x = rng.standard_normal((w["K"], 3)) + 1j * rng.standard_normal((w["K"], 3))
        s = ctx.modulate(wf, x)
        assert rel(s, w["G"] @ x) < 1e-12
        # identity stated at FBMC.m:319-320 / OFDM.m:185-186: G*x(:) == Modulation(x)
        assert rel(s[:, 0], modem.Modulation(x[:, 0].reshape(L, Ksym, order="F"))) < 1e-12
        r = rng.standard_normal((S["N"], 2)) + 1j * rng.standard_normal((S["N"], 2))
        y = ctx.demodulate(wf, r)
        assert rel(y, w["Q"].conj().T @ r) < 1e-12
        assert rel(y[:, 1], modem.Demodulation(r[:, 1]).reshape(-1, order="F")) < 1e-12


def test_k3_estimate(ds_default, gpu_ctx):
    from oracle.ds import _dhat
    S, ctx = ds_default, gpu_ctx
    rng = np.random.default_rng(4)
    for sc in ("aux", "cod", "ofdm"):
        m = S["schemes"][sc]
        w = S["wf"][m["waveform"]]
        for variant, key, isnr in ((0, "W", 0), (1, "W_noInt", len(S["Pn"]) - 1), (0, "W", 3)):
            hP = rng.standard_normal(S["P"]) + 1j * rng.standard_normal(S["P"])
            D_ref, h_ref = _dhat(w, m[key][isnr], hP, faithful=False)
            D, hd = ctx.estimate(sc, variant, isnr, hP)
            assert rel(D, D_ref.toarray()) < TOL
            assert rel(hd, h_ref) < TOL


@pytest.mark.parametrize("n_iter", [0, 1, 4])
def test_loop_body_explicit_draws(ds_default, gpu_ctx, draws3, n_iter):
    """DS.m:350-565 on identical draws: identical bit-error counts, state within 1e-9."""
    from oracle.ds import ds_realization
    import dataclasses
    S, ctx = ds_default, gpu_ctx
    S_it = dict(S)
    S_it["cfg"] = dataclasses.replace(S["cfg"], NrIterations=max(n_iter, 0))
    st, keep = ctx.pack_draws(draws3)
    err = ctx.run_batch(len(draws3), n_iter, st)
    for r, d in enumerate(draws3):
        out = ds_realization(S_it, d, keep=True)
        ref = err_from_oracle(out, n_iter)
        assert np.array_equal(err[r], ref), "bit-error counts differ for realization %d" % r
        for sc in S["schemes"]:
            for isnr in (0, len(S["Pn"]) - 1):
                assert rel(ctx.get_state("y", sc, r, isnr), out["inter"]["y_" + sc][isnr]) < TOL
                assert rel(ctx.get_state("hP", sc, r, isnr), out["inter"]["hP_" + sc][isnr][-1]) < TOL
                assert rel(ctx.get_state("hdiag", sc, r, isnr), out["inter"]["hdiag_" + sc][isnr][-1]) < TOL
                assert rel(ctx.get_state("xD_est", sc, r, isnr), out["inter"]["xD_est_" + sc][isnr][-1]) < 1e-7
                assert rel(ctx.get_state("xD_perf", sc, r, isnr), out["inter"]["xD_perf_" + sc][isnr][-1]) < 1e-7
    nb = ctx.bit_counts()
    for sc, sid in (("aux", 0), ("cod", 1), ("ofdm", 2)):
        m = S["schemes"][sc]
        assert nb[sid, 0] == m["nD"] * m["nbits"] and nb[sid, 1] == int(np.sum(m["considered_bits"]))


def test_batch_is_order_and_size_independent(ds_default, gpu_ctx, draws3):
    """Columns are independent: any batch composition gives the same per-realization counts
    (ragged batch = not a multiple of the 16-column CTA width)."""
    ctx = gpu_ctx
    st, keep = ctx.pack_draws(draws3)
    err = ctx.run_batch(3, 2, st)
    st1, keep1 = ctx.pack_draws([draws3[2]])
    err1 = ctx.run_batch(1, 2, st1)
    assert np.array_equal(err1[0], err[2])
    st2, keep2 = ctx.pack_draws([draws3[i % 3] for i in range(19)])
    err2 = ctx.run_batch(19, 2, st2)
    for i in range(19):
        assert np.array_equal(err2[i], err[i % 3])


def test_seeded_generator_matches_oracle(ds_default, gpu_ctx):
    from oracle import rng
    from oracle.ds import ds_realization
    S, ctx = ds_default, gpu_ctx
    seed, first = 0x1234ABCD5678, 40
    ctx.generate_draws(3, seed, first)
    dev = ctx.download_draws(3)
    for r in range(3):
        ref = rng.draws_for(S, seed, first + r)
        assert np.array_equal(dev[r]["doppler_u"], ref["doppler_u"])
        assert np.array_equal(dev[r]["phase_u"], ref["phase_u"])
        for k in ("bits_aux", "bits_cod", "bits_ofdm", "pil_idx_fbmc", "pil_idx_ofdm"):
            assert np.array_equal(dev[r][k], ref[k]), k
        assert rel(dev[r]["noise"], ref["noise"]) < 1e-12
    # a seeded run equals the oracle run on the generator's own draws, and does not depend on batching
    err = ctx.run_batch(3, 4, None, seed=seed, first_rep=first)
    for r in range(3):
        out = ds_realization(S, dev[r])
        assert np.array_equal(err[r], err_from_oracle(out, 4))
    err_b = ctx.run_batch(1, 4, None, seed=seed, first_rep=first + 2)
    assert np.array_equal(err_b[0], err[2])


def test_round_trip_properties(ds_default, gpu_ctx):
    """Size-independent properties: Q_O' G_O = I, Re(Q_F' G_F) = I (to the Hermite residual),
    linearity of the banded channel operator."""
    S, ctx = ds_default, gpu_ctx
    rng = np.random.default_rng(8)
    Ko, Kf = S["wf"]["O"]["K"], S["wf"]["F"]["K"]
    x = rng.standard_normal((Ko, 2)) + 1j * rng.standard_normal((Ko, 2))
    assert rel(ctx.demodulate("O", ctx.modulate("O", x)), x) < 1e-12
    xr = rng.standard_normal((Kf, 2))
    back = ctx.demodulate("F", ctx.modulate("F", xr))
    assert np.max(np.abs(back.real - xr)) < 1e-5
    ctx.new_realization_seeded(1, 7, 0)
    a = rng.standard_normal(S["N"]) + 1j * rng.standard_normal(S["N"])
    b = rng.standard_normal(S["N"]) + 1j * rng.standard_normal(S["N"])
    lhs = ctx.convolve(2.0 * a - 1j * b, 0)
    assert rel(lhs, 2.0 * ctx.convolve(a, 0) - 1j * ctx.convolve(b, 0)) < 1e-12
