"""Oracle self-tests: the identities the reference states in its comments, its closed-form theory
and hand-derived constants.  The reference ships no golden vectors -- parity unpinned."""
import numpy as np

from oracle.signal_constellation import SignalConstellation
from oracle.fbmc import FBMC
from oracle.ofdm import OFDM
from oracle.bep import bit_error_probability_doubly_flat_rayleigh as bep


def test_geometry_constants(ds_default):
    S = ds_default
    assert S["N"] == 540 and S["P"] == 16
    o = S["ofdm"].Implementation
    assert (o["CyclicPrefix"], o["ZeroGuardSamples"], o["FFTSize"]) == (2, 88, 24)          # CP 1.71 -> 2 (OFDM.m:73-77)
    f = S["fbmc"]
    assert (f.Implementation["TimeSpacing"], f.Nr["SamplesPrototypeFilter"], f.Implementation["FFTSize"]) == (12, 192, 24)
    assert np.allclose(S["chan"].Implementation["PowerDelayProfileNormalized"], [0.9798, 0.0202], atol=5e-5)
    assert abs(S["fD"] - 1158.18) < 0.01
    assert S["wf"]["F"]["K"] == 720 and S["wf"]["O"]["K"] == 336
    assert [S["schemes"][s]["nD"] for s in ("aux", "cod", "ofdm")] == [640, 688, 320]
    assert [S["schemes"][s]["nD"] * S["schemes"][s]["nbits"] for s in ("aux", "cod", "ofdm")] == [2560, 2752, 2560]


def test_stated_identities(ds_default):
    S = ds_default
    rng = np.random.default_rng(0)
    f, o = S["fbmc"], S["ofdm"]
    GF, QF, GO, QO = S["wf"]["F"]["G"], S["wf"]["F"]["Q"], S["wf"]["O"]["G"], S["wf"]["O"]["Q"]
    x = rng.standard_normal((24, 30)) + 1j * rng.standard_normal((24, 30))
    assert np.max(np.abs(GF @ x.reshape(-1, order="F") - f.Modulation(x))) < 1e-12        # FBMC.m:319-320
    r = rng.standard_normal(540) + 1j * rng.standard_normal(540)
    assert np.max(np.abs(QF.conj().T @ r - f.Demodulation(r).reshape(-1, order="F"))) < 1e-12   # FBMC.m:344-345
    xo = rng.standard_normal((24, 14)) + 1j * rng.standard_normal((24, 14))
    assert np.max(np.abs(GO @ xo.reshape(-1, order="F") - o.Modulation(xo))) < 1e-12      # OFDM.m:185-186
    assert np.max(np.abs(QO.conj().T @ r - o.Demodulation(r).reshape(-1, order="F"))) < 1e-12   # OFDM.m:206-207
    assert np.max(np.abs(QO.conj().T @ GO - np.eye(336))) < 1e-14
    QG = QF.conj().T @ GF
    assert np.max(np.abs(QG.real - np.eye(720))) < 5e-7                                    # Hermite residual 4.0e-7
    assert np.max(np.abs(S["D0"] - QG)) < 1e-13                                            # fast GetFBMCMatrix == Q'G
    assert np.max(np.abs(f.GetFBMCMatrix(False) - QG)) < 1e-12                             # FBMC.m:356-357
    w = np.sort(np.unique(np.round(np.abs(S["D0"][:, 0]), 4)))[::-1]
    assert np.allclose(w[:6], [1, 0.4357, 0.2393, 0.0369, 0.0098, 0.0054])


def test_precoder_invariants(ds_default):
    S = ds_default
    a, c = S["aux"], S["cod"]
    assert (a.NrPilotSymbols, a.NrDataSymbols, a.NrAuxiliarySymbols) == (16, 640, 64)
    assert abs(a.PilotToDataPowerOffset * a.DataPowerReduction - S["kappa"]["aux"]) < 1e-15
    assert abs(c.DataPowerReduction - 720 / 752) < 1e-14 and abs(S["kappa"]["cod"] - 4 * 720 / 752) < 1e-13
    assert abs(S["dpr_o"] - 336 / 352) < 1e-15 and abs(S["kappa"]["ofdm"] - 2 * 336 / 352) < 1e-15
    C = c.PrecodingMatrix
    g = C.conj().T @ C
    assert np.allclose(np.diag(g)[16:], c.DataPowerReduction)                              # FBMC.m:595: C'C = I (scaled)
    assert np.max(np.abs(g - np.diag(np.diag(g)))) < 1e-14
    pil = S["wf"]["F"]["pil"]
    for obj in (a, c):                                                                     # IIC.m:9-11
        T = S["D0"][pil, :] @ obj.PrecodingMatrix
        assert np.all(obj.SIR_dB > 25)
        assert np.max(np.abs(T[:, :16].imag)) < 1e-5
    for obj in (a, c):                                                                     # unit mean TX power
        assert abs(np.sum(np.abs(obj.PrecodingMatrix) ** 2) / 720 - 1) < 1e-12
    assert int(S["schemes"]["ofdm"]["considered_bits"].sum()) == 58 * 8


def test_correlation_matrices(ds_default):
    from oracle.ds import ds_setup_literal_check
    S = ds_default
    for wf in ("F", "O"):
        R = S["wf"][wf]["R_hP"]
        assert np.max(np.abs(R - R.conj().T)) < 1e-14 and np.min(np.linalg.eigvalsh(R)) > -1e-12
    assert abs(S["wf"]["F"]["R_hP"][0, 0].real - 0.99502) < 1e-5
    dev = ds_setup_literal_check(S, "ofdm", (0, 9))          # kron / N^2 x N^2 form, FF.m:366-407, DS.m:213-260
    assert max(dev.values()) < 1e-13
    dev = ds_setup_literal_check(S, "cod", (5,))
    assert max(dev.values()) < 1e-13
    for sc in S["schemes"]:
        K = S["wf"][S["schemes"][sc]["waveform"]]["K"]
        frac = np.count_nonzero(S["schemes"][sc]["W"][0]) / (K * K * 16)
        assert (0.15 < frac < 0.18) if sc != "ofdm" else abs(frac - 1 / 14) < 1e-12


def test_constellations():
    for order, method in ((16, "PAM"), (256, "QAM"), (4, "QAM")):
        c = SignalConstellation(order, method)
        assert abs(np.mean(np.abs(c.SymbolMapping) ** 2) - 1) < 1e-14
        nb = c.BitMapping.shape[1]
        assert np.array_equal(c.BitMapping @ (1 << np.arange(nb)), np.arange(order))       # sorted by bit value
        bits = np.random.default_rng(1).integers(0, 2, 200 * nb)
        sym = c.Bit2Symbol(bits)
        assert np.array_equal(c.Symbol2Bit(sym), bits) and np.array_equal(c.SymbolQuantization(sym * 1.01), sym)
        # Gray: nearest neighbours differ in one bit
        d = np.abs(c.SymbolMapping[:, None] - c.SymbolMapping[None, :])
        near = np.isclose(d, np.min(d[d > 1e-9]))
        assert np.all(np.sum(c.BitMapping[:, None, :] != c.BitMapping[None, :, :], axis=2)[near] == 1)


def test_theory_curve_known_values():
    q4 = SignalConstellation(4, "QAM")
    snr = np.array([0.0, 10.0, 20.0])
    closed = 0.5 - 1 / (2 * np.sqrt(2 * (1 + 10 ** (-snr / 10)) - 1))                     # SV.m:179 (4-QAM closed form)
    assert np.max(np.abs(bep(snr, q4.SymbolMapping, q4.BitMapping) - closed)) < 1e-12
    q256 = SignalConstellation(256, "QAM")
    v = bep([32.0], q256.SymbolMapping, q256.BitMapping)[0]
    assert abs(v - 0.0130) < 5e-4                                                          # png/Figure5.png grey line


def test_doubly_flat_simulation_matches_theory():
    """SV.m:118-169 (perfect-CSI branch) with the FFT modem: y/h demapped against the theory curve."""
    rng = np.random.default_rng(3)
    o = OFDM(12, 15, 15e3, 15e3 * 14 * 12, 15e3 * 20, False, 0, (8 - 1 / 2) * 1 / 15e3 * 1 / 2)
    qam = SignalConstellation(16, "QAM")
    snr_db, errs, tot = 15.0, 0, 0
    Pn = o.PHY["SamplingRate"] / (o.PHY["SubcarrierSpacing"] * o.Nr["Subcarriers"]) * 10 ** (-snr_db / 10)   # SV.m:92
    for _ in range(400):
        bits = rng.integers(0, 2, 12 * 15 * 4)
        x = qam.Bit2Symbol(bits).reshape(12, 15, order="F")
        h = np.sqrt(0.5) * (rng.standard_normal() + 1j * rng.standard_normal())
        s = o.Modulation(x)
        n = np.sqrt(Pn / 2) * (rng.standard_normal(len(s)) + 1j * rng.standard_normal(len(s)))
        y = o.Demodulation(h * s + n)
        errs += np.sum(qam.Symbol2Bit(y.reshape(-1, order="F") / h) != bits)
        tot += len(bits)
    theory = bep([snr_db], qam.SymbolMapping, qam.BitMapping)[0]
    assert abs(errs / tot - theory) < 0.25 * theory


def _bundle_draws(B, S, reps):
    """Rebuild per-repetition draws from the three replay streams of a reference bundle (consumption order of
    FF.m:227-233, DS.m:355-367, DS.m:399 -- see oracle/export_reference_bundle.py)."""
    u, i, g = (np.asarray(B["stream_" + k]).reshape(-1) for k in ("rand", "randi", "randn"))
    T, paths, N, nS = len(S["chan"].Implementation["IndexDelayTaps"]), S["cfg"].Paths, S["N"], len(S["Pn"])
    nb = {sc: len(S["schemes"][sc]["considered_bits"]) for sc in ("aux", "cod", "ofdm")}
    P = S["P"]
    out, pu, pi, pg = [], 0, 0, 0
    for _ in range(reps):
        d = {}
        d["doppler_u"] = u[pu:pu + T * paths].reshape(T, paths, order="F"); pu += T * paths
        d["phase_u"] = u[pu:pu + T * paths].reshape(T, paths, order="F"); pu += T * paths
        for sc in ("aux", "cod", "ofdm"):
            d["bits_" + sc] = i[pi:pi + nb[sc]].astype(np.uint8); pi += nb[sc]
        d["pil_idx_fbmc"] = i[pi:pi + P].astype(np.int64) - 1; pi += P
        d["pil_idx_ofdm"] = i[pi:pi + P].astype(np.int64) - 1; pi += P
        nz = np.zeros((nS, N), dtype=complex)
        for s_ in range(nS):
            nz[s_] = g[pg:pg + N] + 1j * g[pg + N:pg + 2 * N]; pg += 2 * N
        d["noise"] = nz
        out.append(d)
    assert (pu, pi, pg) == (len(u), len(i), len(g))            # the streams are consumed exactly
    return out


def test_reference_bundle_default_replays(ds_default):
    """tests/golden/reference_bundle_default.mat (for matlab/verify_oracle.m): the committed replay streams, read back in
    the reference's consumption order, reproduce the committed 24 BER arrays and end-of-script vectors."""
    import os
    import scipy.io
    from oracle.ds import ds_realization, ber_arrays
    S = ds_default
    B = scipy.io.loadmat(os.path.join(os.path.dirname(__file__), "golden", "reference_bundle_default.mat"),
                         squeeze_me=True, struct_as_record=False)
    reps = int(B["overrides"].NrRepetitions)
    assert list(B["overrides"].M_SNR_dB) == list(S["cfg"].M_SNR_dB) and B["overrides"].paper_block == 0
    draws = _bundle_draws(B, S, reps)
    res = [ds_realization(S, d, keep=(r == reps - 1)) for r, d in enumerate(draws)]
    ber = ber_arrays(res, S)
    assert len(ber) == 24
    for name, arr in ber.items():
        assert np.array_equal(arr, getattr(B["ber"], name).reshape(arr.shape)), name
    last = res[-1]["inter"]
    assert np.max(np.abs(last["hP_aux"][-1][-1] - B["last"].hP_est_FBMC_Aux_Temp)) < 1e-12
    assert np.max(np.abs(np.diag(last["D_O"]) - B["last"].h_OFDM)) < 1e-12
    assert np.max(np.abs(S["cod"].SIR_dB - B["setup"].SIR_dB_Cod)) < 1e-9


def test_factored_estimate_equals_thresholded_weights_up_to_the_thresholds():
    """D-hat = sum_p W_p hP(p) (DS.m:417-425) against its factored form Q^H (sum_q g_q M_q) G, g = pinv(R_hP_est) hP (what the
    library's factored estimator applies): for CP-OFDM the 1e-8 thresholds of DS.m:263-264, 287-289 only remove rounding noise, so
    the two agree to rounding; for FBMC they differ by the removed entries (a few 1e-5 of max|D-hat| at high SNR)."""
    from oracle.ds import DSConfig, ds_setup, _dhat, _dhat_factored, _offdiag_times
    S = ds_setup(DSConfig())
    rng_ = np.random.default_rng(5)
    for sc, tol_lo, tol_hi in (("ofdm", 0.0, 1e-11), ("aux", 1e-9, 1e-4)):
        m = S["schemes"][sc]
        w = S["wf"][m["waveform"]]
        K, P = w["K"], S["P"]
        for isnr in (0, len(S["Pn"]) - 1):
            hP = (rng_.standard_normal(P) + 1j * rng_.standard_normal(P)) / np.sqrt(2)
            v = (rng_.standard_normal(K) + 1j * rng_.standard_normal(K)) / np.sqrt(2)
            Dh, hh = _dhat(w, m["W"][isnr], hP, False)
            Df, hf = _dhat_factored(w, m["W"][isnr], m["Rinv"][isnr], hP)
            assert np.array_equal(hh, hf)
            a, b = _offdiag_times(Dh, hh, v), _offdiag_times(Df, hf, v)
            dev = np.max(np.abs(a - b)) / np.max(np.abs(a))
            assert tol_lo <= dev < tol_hi, (sc, isnr, dev)
