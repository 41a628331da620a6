"""Oracle restatement of DoublySelectiveChannelEstimation.m: the setup (DS.m:16-313) and the
Monte-Carlo loop body (DS.m:350-565).  Test infrastructure only -- see oracle/__init__.py.

The correlation matrices are computed in structured form (never building the N^2 x N^2
R_vecH): with H[a+m, a] = h[a+m, m], E{H[pos_m(a)] conj(H[pos_m(b)])} = PDP_m R_t((a-b) dt)
(FastFading.m:372-404), hence
    R_vecH * vec(q g^H)  ->  banded pseudo-channel  M[pos_m(a)] = PDP_m (R_t * zeta_m)[a],
    zeta_m[b] = q[row_m(b)] conj(g[col_m(b)])
and R_Dij_hP(:,p) = vec(Q^H M_p G) (DS.m:260).  pos_m() reproduces the reference's linear
index arithmetic including its wrap of the last columns (FastFading.m:377,406).
``ds_setup_literal_check`` rebuilds selected quantities literally with scipy.sparse / kron for
cross-checking.
"""
from dataclasses import dataclass, field
import numpy as np
import scipy.sparse as sp

from .fbmc import FBMC
from .ofdm import OFDM, _matlab_round
from .signal_constellation import SignalConstellation
from .fast_fading import FastFading
from .iic import ImaginaryInterferenceCancellationAtPilotPosition as IIC

SCHEMES = ("aux", "cod", "ofdm")


@dataclass
class DSConfig:
    """DS.m:16-37 (defaults) -- ``paper()`` applies DS.m:42-46."""
    M_SNR_dB: tuple = tuple(range(10, 41, 5))
    NrRepetitions: int = 25
    ZeroThresholdSparse: float = 8
    L: int = 24
    F: float = 15e3
    SamplingRate: float = 15e3 * 24
    NrSubframes: int = 1
    QAM_ModulationOrder: int = 256
    PilotToDataPowerOffset: float = 2
    PilotToDataPowerOffsetAux: float = 4.685
    NrIterations: int = 4
    Velocity_kmh: float = 500
    PowerDelayProfile: object = "VehicularA"
    DopplerModel: str = "Jakes"
    Paths: int = 200
    schemes: tuple = SCHEMES

    @staticmethod
    def paper(**kw):
        base = dict(M_SNR_dB=tuple(range(10, 41, 2)), NrRepetitions=1000, SamplingRate=15e3 * 14 * 14, NrSubframes=2)
        base.update(kw)
        return DSConfig(**base)


def pilot_matrices(L, NrSubframes):
    """DS.m:91-113 (MATLAB 1-based strided ranges restated 0-based)."""
    pm_o = np.zeros((L, 14))
    for r0, c0 in ((1, 1), (4, 5), (7, 1), (10, 5)):          # rows 2,5,8,11 ; cols 2:7:end / 6:7:end
        pm_o[r0::12, c0::7] = 1
    pm_o = np.tile(pm_o, (1, NrSubframes))
    pm_f = np.zeros((L, 30))
    for r0, c0 in ((1, 2), (4, 10), (7, 3), (10, 11)):        # cols 3:16:end, 11:16:end, 4:16:end, 12:16:end
        pm_f[r0::12, c0::16] = 1
    pm_f = np.tile(pm_f, (1, NrSubframes))
    aux = pm_f.copy()
    a, b = np.nonzero(pm_f)
    for i in range(len(a)):                                   # :108-113
        aux[a[i] + 1, b[i]] = -1
        aux[a[i] - 1, b[i]] = -1
        aux[a[i], b[i] + 1] = -1
        aux[a[i], b[i] - 1] = -1
    return pm_o, pm_f, aux


def _positions(N, m):
    """Linear-index arithmetic of FastFading.m:377: entry (a, m) of IndexCorrMatrixConv addresses
    H(:)[a*(N+1)+m]; returns (row, col, valid)."""
    a = np.arange(N)
    lin = a * (N + 1) + m
    valid = lin < N * N
    lin = np.where(valid, lin, 0)
    return lin % N, lin // N, valid


class _Corr:
    """Structured access to R_vecH for one channel model."""

    def __init__(self, chan):
        self.N = chan.Nr["SamplesTotal"]
        self.pdp = chan.Implementation["PowerDelayProfileNormalized"]
        rt, _ = chan.GetTimeCorrelation()
        self.rt = rt                                           # index N-1 <-> lag 0

    def toep(self, rows, cols):
        return self.rt[(self.N - 1) + rows[:, None] - cols[None, :]]

    def pseudo_channel(self, q, g):
        """reshape(R_vecH*kron(g.',q')',N,N) of DS.m:213,260 as scipy CSR."""
        N = self.N
        rr, cc, vv = [], [], []
        for m in np.flatnonzero(self.pdp):
            row, col, valid = _positions(N, m)
            zeta = np.where(valid, q[row] * np.conj(g[col]), 0)
            nz = np.flatnonzero(zeta)
            mu = self.pdp[m] * (self.toep(np.arange(N), nz) @ zeta[nz])
            rr.append(row[valid]); cc.append(col[valid]); vv.append(mu[valid])
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(N, N))

    def pilot_power(self, q, GA, kappa):
        """abs(sum(sum((GA.'*(Temp*R_vecH*Temp')).*GA',2))) of DS.m:224-233."""
        N = self.N
        tot = 0.0
        for m in np.flatnonzero(self.pdp):
            row, col, valid = _positions(N, m)
            sel = np.flatnonzero(valid & (q[row] != 0))
            U = np.conj(q[row[sel]])[:, None] * GA[col[sel], :]
            tot = tot + self.pdp[m] / kappa * np.sum(np.conj(U) * (self.toep(sel, sel) @ U))
        return abs(tot)


def ds_setup(cfg: DSConfig, verbose=False):
    """DS.m:50-313.  Returns a dict with everything the loop body needs."""
    S = {}
    L, F, fs = cfg.L, cfg.F, cfg.SamplingRate
    fbmc = FBMC(L, 30 * cfg.NrSubframes, F, fs, 0, False, "Hermite-OQAM", 8, 0, True)       # :51-62
    zg = ((fbmc.Nr["SamplesTotal"] - (_matlab_round((1 / 15e3 / 14) * fs) + _matlab_round(fs / 15e3))
           * 14 * cfg.NrSubframes) / 2) / fs                                                # :66
    ofdm = OFDM(L, 14 * cfg.NrSubframes, F, fs, 0, False, 1 / 15e3 / 14, zg)                # :67-76
    if ofdm.Nr["SamplesTotal"] != fbmc.Nr["SamplesTotal"]:                                  # :79-81
        raise ValueError("Total number of samples must be the same for OFDM and FBMC.")
    N = ofdm.Nr["SamplesTotal"]
    PAM = SignalConstellation(int(round(np.sqrt(cfg.QAM_ModulationOrder))), "PAM")          # :86
    QAM = SignalConstellation(cfg.QAM_ModulationOrder, "QAM")                               # :87
    pm_o, pm_f, pm_aux = pilot_matrices(L, cfg.NrSubframes)                                 # :91-113
    D0 = fbmc.GetFBMCMatrix()
    aux = IIC("Auxiliary", pm_aux, D0, 28, cfg.PilotToDataPowerOffsetAux)                   # :116-122
    cod = IIC("Coding", pm_f, D0, 20, 2 * cfg.PilotToDataPowerOffset)                       # :123-129
    pmo = pm_o.reshape(-1, order="F")
    pmf = pm_f.reshape(-1, order="F")
    pmaux = pm_aux.reshape(-1, order="F")
    P = int(np.sum(pmo == 1))                                                               # :131
    nD_o = int(np.sum(pmo == 0))                                                            # :132
    Ko = pmo.size
    map_o = np.zeros((Ko, Ko))                                                              # :134-137
    map_o[np.flatnonzero(pmo == 1), np.arange(P)] = np.sqrt(cfg.PilotToDataPowerOffset)
    map_o[np.flatnonzero(pmo == 0), P + np.arange(nD_o)] = 1.0
    map_o = map_o / np.sqrt(np.mean(np.sum(map_o**2, axis=1)))
    dpr_o = Ko / (P * cfg.PilotToDataPowerOffset + nD_o)                                    # :138
    kappa = {"aux": aux.PilotToDataPowerOffset * aux.DataPowerReduction,                    # :140-142
             "cod": cod.PilotToDataPowerOffset * cod.DataPowerReduction,
             "ofdm": cfg.PilotToDataPowerOffset * dpr_o}
    # ---- "no edge" masks, DS.m:145-172
    ct_f = np.zeros(pm_f.shape); ct_f[4:-4, 10:-10] = 1
    ct_o = np.zeros(pm_o.shape); ct_o[4:-4, 5:-5] = 1
    ctf, cto = ct_f.reshape(-1, order="F"), ct_o.reshape(-1, order="F")
    Ca, Cc = aux.PrecodingMatrix, cod.PrecodingMatrix
    m_aux = (ctf * (pmaux == 0)) == 1
    cons_aux = np.sum(np.abs(Ca[m_aux, P:]), axis=0) > aux.DataPowerReduction * 0.9         # :155
    cons_cod = ~np.any(Cc[ctf == 0, P:] != 0, axis=0)                                       # :161
    m_o = (cto * (pmo == 0)) == 1
    cons_o = np.sum(np.abs(map_o[m_o, P:]), axis=0) > dpr_o * 0.9                           # :167
    nb_pam = PAM.BitMapping.shape[1]
    nb_qam = QAM.BitMapping.shape[1]
    cbits = {"aux": np.repeat(cons_aux, nb_pam), "cod": np.repeat(cons_cod, nb_pam),        # :170-172
             "ofdm": np.repeat(cons_o, nb_qam)}
    # ---- channel + TX/RX matrices, DS.m:176-205
    fD = cfg.Velocity_kmh / 3.6 * 2.5e9 / 2.998e8
    chan = FastFading(fs, cfg.PowerDelayProfile, N, fD, cfg.DopplerModel, cfg.Paths, 1, 1, False)
    G_F = fbmc.GetTXMatrix(); Q_F = fbmc.GetRXMatrix().conj().T                             # :191-192
    G_O = ofdm.GetTXMatrix(); Q_O = ofdm.GetRXMatrix().conj().T                             # :194-195
    pil_f, pil_o = np.flatnonzero(pmf == 1), np.flatnonzero(pmo == 1)
    G_pre = {"aux": G_F @ Ca, "cod": G_F @ Cc, "ofdm": G_O @ map_o}                         # :203-205
    corr = _Corr(chan)
    Pn = [fs / (F * L) * 10 ** (-snr / 10) for snr in cfg.M_SNR_dB]                         # :243
    thr = 10.0 ** (-cfg.ZeroThresholdSparse)
    wf = {"F": dict(G=G_F, Q=Q_F, pil=pil_f, K=G_F.shape[1]),
          "O": dict(G=G_O, Q=Q_O, pil=pil_o, K=G_O.shape[1])}
    for name, w in wf.items():                                                              # :208-215, 256-268
        if verbose:
            print("correlation matrices, waveform", name)
        K = w["K"]
        R_hP = np.zeros((P, P), dtype=complex)
        cols = []
        w["Mq"] = []                                           # pseudo-channels, kept for the factored estimator (ds_realization)
        for jp in range(P):
            M = corr.pseudo_channel(w["Q"][:, w["pil"][jp]], w["G"][:, w["pil"][jp]])
            w["Mq"].append(M)
            Dj = w["Q"].conj().T @ (M @ w["G"])                                             # :260
            R_hP[:, jp] = Dj[w["pil"], w["pil"]]                                            # :213
            v = Dj.reshape(-1, order="F")
            v[np.abs(v) < thr] = 0                                                          # :263
            cols.append(sp.csc_matrix(v.reshape(-1, 1)))
        R_D = sp.hstack(cols).tocsr()                                                       # K^2 x P
        sup = np.flatnonzero(np.diff(R_D.indptr))                                           # rows with any non-zero
        w["R_hP"], w["sup"], w["R_sup"] = R_hP, sup, R_D[sup, :].toarray()
    schemes = {}
    spec = {"aux": ("F", aux, Ca), "cod": ("F", cod, Cc), "ofdm": ("O", None, map_o.astype(complex))}
    for sc in cfg.schemes:
        wname, obj, C = spec[sc]
        w = wf[wname]
        R_hP = w["R_hP"]
        R_nn = R_hP.copy()                                                                  # :219-234
        qn = np.zeros(P)
        for ip in range(P):
            q = w["Q"][:, w["pil"][ip]]
            R_nn[ip, ip] = corr.pilot_power(q, G_pre[sc], kappa[sc])
            qn[ip] = np.real(np.vdot(q, q))
        W, W_ni, Ri, Ri_ni = [], [], [], []
        for isnr in range(len(Pn)):                                                         # :238-313
            R_est = R_nn.copy()
            R_est[np.arange(P), np.arange(P)] = np.diag(R_nn) + Pn[isnr] * qn / kappa[sc]   # :245-247
            R_ni = R_est - (R_nn - R_hP)                                                    # :251-253
            for R, out, inv in ((R_est, W, Ri), (R_ni, W_ni, Ri_ni)):
                inv.append(np.linalg.pinv(R))
                Wv = w["R_sup"] @ inv[-1]                                                   # :283 / :302
                Wv[np.abs(Wv) < thr] = 0                                                    # :287 / :306
                out.append(Wv)
        schemes[sc] = dict(waveform=wname, C=C, W=W, W_noInt=W_ni, Rinv=Ri, Rinv_noInt=Ri_ni, kappa=kappa[sc],
                           R_hP_est_noNoise=R_nn, considered_bits=cbits[sc])
    schemes_meta = {
        "aux": dict(nD=aux.NrDataSymbols, data_idx=np.flatnonzero(pmaux == 0), dpr=aux.DataPowerReduction,
                    const="PAM", nbits=nb_pam),
        "cod": dict(nD=cod.NrDataSymbols, data_idx=None, dpr=cod.DataPowerReduction, const="PAM", nbits=nb_pam),
        "ofdm": dict(nD=nD_o, data_idx=np.flatnonzero(pmo == 0), dpr=dpr_o, const="QAM", nbits=nb_qam),
    }
    for sc in schemes:
        schemes[sc].update(schemes_meta[sc])
    S.update(cfg=cfg, N=N, P=P, fbmc=fbmc, ofdm=ofdm, PAM=PAM, QAM=QAM, chan=chan, aux=aux, cod=cod,
             pm_o=pm_o, pm_f=pm_f, pm_aux=pm_aux, map_o=map_o, dpr_o=dpr_o, kappa=kappa,
             wf=wf, schemes=schemes, Pn=np.array(Pn), fD=fD, D0=D0)
    for name, w in wf.items():                                                              # CSR pattern of D-hat
        K = w["K"]
        i, j = w["sup"] % K, w["sup"] // K
        order = np.lexsort((j, i))
        indptr = np.concatenate([[0], np.cumsum(np.bincount(i, minlength=K))])
        w["csr"] = (order, j[order], indptr)
        w["diag_pos"] = np.searchsorted(w["sup"], np.arange(K) * (K + 1))                   # position of (i,i)
        assert np.array_equal(w["sup"][w["diag_pos"]], np.arange(K) * (K + 1))
    return S


def new_draws(S, rng):
    """One realization's random draws in the order DS.m:352-368,399 consumes them."""
    cfg = S["cfg"]
    T = len(S["chan"].Implementation["IndexDelayTaps"])
    if S["chan"].Implementation.get("UseDiscreteDopplerSpectrum"):         # FF.m:208-209: normals, not uniforms
        nb = S["chan"].Implementation["DiscreteDopplerSpectrum"].shape[0]
        d = dict(gauss=rng.standard_normal((nb, T)) + 1j * rng.standard_normal((nb, T)))
    else:
        d = dict(doppler_u=rng.random((T, cfg.Paths)), phase_u=rng.random((T, cfg.Paths)))
    for sc in ("aux", "cod", "ofdm"):
        if sc in S["schemes"]:
            m = S["schemes"][sc]
            d["bits_" + sc] = rng.integers(0, 2, m["nD"] * m["nbits"]).astype(np.uint8)
    d["pil_idx_fbmc"] = rng.integers(0, S["PAM"].ModulationOrder, S["P"]).astype(np.int32)
    d["pil_idx_ofdm"] = rng.integers(0, S["QAM"].ModulationOrder, S["P"]).astype(np.int32)
    d["noise"] = rng.standard_normal((len(cfg.M_SNR_dB), S["N"])) + 1j * rng.standard_normal(
        (len(cfg.M_SNR_dB), S["N"]))
    return d


def _dhat(w, Wv, hP, faithful):
    """DS.m:417-425: D-hat = sum_p W(:,:,p) hP(p), returned as scipy CSR (+ its diagonal)."""
    K = w["K"]
    if faithful:       # densify + broadcast multiply + sum over pilots, as the reference does
        full = np.zeros((K * K, Wv.shape[1]), dtype=complex)
        full[w["sup"], :] = Wv
        D = np.sum(full.reshape(K, K, -1, order="F") * hP.reshape(1, 1, -1), axis=2)
        return D, np.diag(D).copy()
    vals = Wv @ hP
    order, idx, indptr = w["csr"]
    return sp.csr_matrix((vals[order], idx, indptr), shape=(K, K)), vals[w["diag_pos"]]


def _dhat_factored(w, Wv, Rinv, hP):
    """The same estimate in factored form: D-hat = Q^H H-hat G with the estimated channel H-hat = sum_q g_q M_q,
    g = pinv(R_hP_est) hP -- DS.m:417-425 with W = R_Dij_hP pinv(R) (DS.m:283) and R_Dij_hP(:, q) = vec(Q^H M_q G)
    (DS.m:260) substituted, WITHOUT the two 1e-8 thresholds (DS.m:263-264, 287-289).  The diagonal h-hat stays
    W_diag hP (what the equaliser uses).  Checks the library's CHEST_ESTIMATOR_FACTORED mode; not the reference's values."""
    g = Rinv @ hP
    Hh = w["Mq"][0] * g[0]
    for q in range(1, len(g)):
        Hh = Hh + w["Mq"][q] * g[q]
    return ("factored", w, Hh.tocsr()), Wv[w["diag_pos"]] @ hP


def pseudo_channel_taps(S, wname):
    """M[p, tap, n] = M_p[n, n - delay_tap] over the non-zero taps of the power delay profile (chest_set_pseudo_channels)."""
    w, N = S["wf"][wname], S["N"]
    delays = np.flatnonzero(S["chan"].Implementation["PowerDelayProfileNormalized"])
    M = np.zeros((len(w["Mq"]), len(delays), N), dtype=complex)
    for p, Mq in enumerate(w["Mq"]):
        for t, d in enumerate(delays):
            M[p, t, d:] = Mq.diagonal(-int(d))
    return M


def _offdiag_times(D, h, v):
    """(D - diag(h)) * v, DS.m:482-484 (h is the diagonal of D, so the diagonal is exactly 0)."""
    if isinstance(D, tuple):                                   # factored estimate
        _, w, Hh = D
        return w["Q"].conj().T @ (Hh @ (w["G"] @ v)) - h * v
    if sp.issparse(D):
        Doff = D.copy()
        Doff.setdiag(0)
        return Doff @ v
    return (D - np.diag(h)) @ v


def ds_realization(S, draws, faithful=False, keep=False, factored=()):
    """DS.m:352-563 for one i_rep.  Returns {'err': {...}, 'nbits': {...}} with integer bit-error
    counts err[scheme][csi][edge] of shape (n_SNR, 1+NrIterations) (column 0 = one-tap equaliser)
    and, with keep=True, the intermediates used as golden vectors.  `factored`: schemes whose D-hat of the cancellation is
    formed as Q^H H-hat G (_dhat_factored) -- the restatement of the library's factored estimator, not of the reference."""
    cfg, P, N = S["cfg"], S["P"], S["N"]
    nS, nI = len(cfg.M_SNR_dB), cfg.NrIterations
    chan = S["chan"]
    chan.NewRealization(draws.get("doppler_u"), draws.get("phase_u"), draws.get("gauss"))   # :352
    H = chan.GetConvolutionMatrix()                                                         # :381
    const = {"PAM": S["PAM"], "QAM": S["QAM"]}
    out = {"err": {}, "nbits": {}, "nbits_noedge": {}}
    inter = {"h": chan.ImpulseResponse.copy()} if keep else None
    Dw, hw = {}, {}
    for wname in {S["schemes"][sc]["waveform"] for sc in S["schemes"]}:
        w = S["wf"][wname]
        Dw[wname] = w["Q"].conj().T @ (H @ w["G"])                                          # :388-389
        hw[wname] = np.diag(Dw[wname]).copy()                                               # :392-393
        if keep:
            inter["D_" + wname] = Dw[wname]
    st = {}
    for sc, m in S["schemes"].items():
        cst = const[m["const"]]
        bits = draws["bits_" + sc]
        xD = cst.Bit2Symbol(bits)                                                           # :360-362
        pidx = draws["pil_idx_fbmc"] if m["waveform"] == "F" else draws["pil_idx_ofdm"]
        xP = cst.SymbolMapping[pidx]                                                        # :365-368
        xP = xP / np.abs(xP)
        w = S["wf"][m["waveform"]]
        x = m["C"] @ np.concatenate([xP, xD])                                               # :371-373
        s = w["G"] @ x                                                                      # :376-378
        r0 = H @ s                                                                          # :383-385
        st[sc] = dict(cst=cst, bits=bits, xP=xP, r0=r0, w=w)
        out["err"][sc] = {c: {e: np.zeros((nS, 1 + nI), dtype=np.int64) for e in ("all", "noedge")}
                          for c in ("est", "perfect")}
        out["nbits"][sc] = len(bits)
        out["nbits_noedge"][sc] = int(np.sum(m["considered_bits"]))
        if keep:
            inter["x_" + sc], inter["r0_" + sc] = x, r0
            inter.update({k + sc: [] for k in ("y_", "hP_", "xD_est_", "xD_perf_", "hdiag_")})

    def detect(sc, x_eq):
        """data selection + scaling + demap + error count, DS.m:430-433 etc."""
        m, s_ = S["schemes"][sc], st[sc]
        if sc == "aux":
            xD_est = np.real(x_eq[m["data_idx"]] / np.sqrt(m["dpr"]))                       # :430
        elif sc == "cod":
            xD_est = np.real((m["C"].conj().T @ x_eq)[P:]) / m["dpr"]                       # :436-437
        else:
            xD_est = x_eq[m["data_idx"]] / np.sqrt(m["dpr"])                                # :444
        e = s_["bits"] != s_["cst"].Symbol2Bit(xD_est)
        return xD_est, int(e.sum()), int(e[m["considered_bits"]].sum())

    for isnr in range(nS):                                                                  # :395
        noise = np.sqrt(S["Pn"][isnr] / 2) * draws["noise"][isnr]                           # :399
        for sc, m in S["schemes"].items():
            s_ = st[sc]
            w, cst, xP = s_["w"], s_["cst"], s_["xP"]
            err = out["err"][sc]
            y = w["Q"].conj().T @ (s_["r0"] + noise)                                        # :401-409
            hP = y[w["pil"]] / xP / np.sqrt(m["kappa"])                                     # :412-414
            if sc in factored:
                Dh, hh = _dhat_factored(w, m["W"][isnr], m["Rinv"][isnr], hP)
            else:
                Dh, hh = _dhat(w, m["W"][isnr], hP, faithful)                               # :417-428
            xD_e, e1, e2 = detect(sc, y / hh)                                               # :429-447
            err["est"]["all"][isnr, 0], err["est"]["noedge"][isnr, 0] = e1, e2
            D, h = Dw[m["waveform"]], hw[m["waveform"]]
            xD_p, e1, e2 = detect(sc, y / h)                                                # :450-466
            err["perfect"]["all"][isnr, 0], err["perfect"]["noedge"][isnr, 0] = e1, e2
            if keep:
                inter["y_" + sc].append(y); inter["hP_" + sc].append([hP])
                inter["xD_est_" + sc].append([xD_e]); inter["xD_perf_" + sc].append([xD_p])
                inter["hdiag_" + sc].append([hh])
                if isnr == 0 and not isinstance(Dh, tuple):
                    inter["Dhat0_" + sc] = Dh.toarray() if sp.issparse(Dh) else Dh
            for it in range(1, nI + 1):                                                     # :481
                v = m["C"] @ np.concatenate([xP, cst.SymbolQuantization(xD_e)])
                y_ic = y - _offdiag_times(Dh, hh, v)                                        # :482-484
                hP = y_ic[w["pil"]] / xP / np.sqrt(m["kappa"])                              # :487-489
                Wsel = m["W"] if it <= nI / 2 else m["W_noInt"]                             # :492
                if sc in factored:
                    Dh, hh = _dhat_factored(w, Wsel[isnr], (m["Rinv"] if it <= nI / 2 else m["Rinv_noInt"])[isnr], hP)
                else:
                    Dh, hh = _dhat(w, Wsel[isnr], hP, faithful)                             # :493-517
                xD_e, e1, e2 = detect(sc, y_ic / hh)                                        # :519-537
                err["est"]["all"][isnr, it], err["est"]["noedge"][isnr, it] = e1, e2
                vp = m["C"] @ np.concatenate([xP, cst.SymbolQuantization(xD_p)])
                y_icp = y - _offdiag_times(D, h, vp)                                        # :541-543
                xD_p, e1, e2 = detect(sc, y_icp / h)                                        # :545-561
                err["perfect"]["all"][isnr, it], err["perfect"]["noedge"][isnr, it] = e1, e2
                if keep:
                    inter["hP_" + sc][-1].append(hP); inter["xD_est_" + sc][-1].append(xD_e)
                    inter["xD_perf_" + sc][-1].append(xD_p); inter["hdiag_" + sc][-1].append(hh)
    if keep:
        out["inter"] = inter
    return out


def ber_arrays(results, S):
    """Assemble the 24 arrays of DS.m:322-345 (S x reps [x I]) from a list of ds_realization
    outputs; names follow the reference's variables."""
    names = {"aux": "FBMC_Aux", "cod": "FBMC_Cod", "ofdm": "OFDM"}
    nS, nI = len(S["cfg"].M_SNR_dB), S["cfg"].NrIterations
    R = len(results)
    arr = {}
    for sc in S["schemes"]:
        for csi, ctag in (("est", ""), ("perfect", "_PerfectCSI")):
            for edge, etag in (("all", ""), ("noedge", "_NoEdge")):
                nb = results[0]["nbits"][sc] if edge == "all" else results[0]["nbits_noedge"][sc]
                e = np.stack([r["err"][sc][csi][edge] for r in results], axis=1) / nb      # S x R x (1+I)
                arr["BER_%s_OneTapEqualizer%s%s" % (names[sc], ctag, etag)] = e[:, :, 0]
                if csi == "est":
                    arr["BER_%s_InterferenceCancellation%s" % (names[sc], etag)] = e[:, :, 1:]
                else:
                    arr["BER_%s_PerfectCSI_InterferenceCancellation%s" % (names[sc], etag)] = e[:, :, 1:]
    return arr


def ds_setup_literal_check(S, scheme="ofdm", pilots=(0, 1)):
    """Rebuild R_vecH literally (FastFading.m:366-407) and evaluate DS.m:213, 224-233 and 260 with
    the reference's kron/reshape expressions for a few pilots.  Returns max abs deviations from
    the structured values in S.  Feasible for N ~ 500 (R_vecH has ~T*N^2 non-zeros)."""
    N, P = S["N"], S["P"]
    m = S["schemes"][scheme]
    w = S["wf"][m["waveform"]]
    R = S["chan"].GetCorrelationMatrix()
    G, Q, pil = w["G"], w["Q"], w["pil"]
    GA = G @ m["C"]
    dev = {"R_hP": 0.0, "R_Dij_hP": 0.0, "R_hP_est_diag": 0.0}
    K = w["K"]
    for jp in pilots:
        z = np.kron(G[:, pil[jp]], Q[:, pil[jp]].conj()).conj()                             # kron(g.',q')'
        Mj = (R @ z).reshape(N, N, order="F")
        col = np.sum((Q[:, pil].conj().T @ Mj) * G[:, pil].T, axis=1)                       # :213
        dev["R_hP"] = max(dev["R_hP"], np.max(np.abs(col - w["R_hP"][:, jp])))
        Dj = (Q.conj().T @ Mj @ G).reshape(-1, order="F")                                   # :260
        Dj[np.abs(Dj) < 10.0 ** (-S["cfg"].ZeroThresholdSparse)] = 0
        ref = np.zeros(K * K, dtype=complex)
        ref[w["sup"]] = w["R_sup"][:, jp]
        dev["R_Dij_hP"] = max(dev["R_Dij_hP"], np.max(np.abs(Dj - ref)))
        T = sp.kron(sp.identity(N, format="csr"), sp.csr_matrix(Q[:, pil[jp]].conj().reshape(1, -1))) \
            / np.sqrt(m["kappa"])                                                           # :224
        T2 = (T @ R @ T.conj().T).toarray()
        val = abs(np.sum((GA.T @ T2) * GA.conj().T))                                        # :225
        dev["R_hP_est_diag"] = max(dev["R_hP_est_diag"], abs(val - m["R_hP_est_noNoise"][jp, jp].real))
    return dev
