"""Tier-2 host driver: DoublySelectiveChannelEstimation.m with its loop body (DS.m:350-565) replaced
by batched device launches.

`DoublySelectiveSimulation(**parameters)` keeps the script's parameter block (DS.m:16-37), runs the
one-time setup (DS.m:50-313: modem objects, pilot patterns, precoders, correlation matrices, MMSE
matrices) and hands its outputs to the CUDA library; `run()` returns the 24 BER arrays of
DS.m:322-345.  The setup's dominant contraction  R_Dij_hP(:,p) = vec(Q' M_p G)  (DS.m:260) is the
same banded-operator GEMM as the hot path's D = Q' H G, so it runs through K2 on the GPU with the
pseudo-channel M_p uploaded as an impulse response; everything else in the setup is small host
linear algebra, as in the reference."""
import time

import numpy as np
import scipy.sparse as sp

from .channel import FastFading
from .context import DeviceContext, SCHEME_ID
from .estimation import ImaginaryInterferenceCancellationAtPilotPosition as _IIC
from .modulation import FBMC, OFDM, SignalConstellation, _round_half_away

BER_NAMES = {"aux": "FBMC_Aux", "cod": "FBMC_Cod", "ofdm": "OFDM"}


class DoublySelectiveSimulation:
    def __init__(self, M_SNR_dB=tuple(range(10, 41, 5)), NrRepetitions=25, ZeroThresholdSparse=8,
                 L=24, F=15e3, SamplingRate=15e3 * 24, NrSubframes=1, QAM_ModulationOrder=256,
                 PilotToDataPowerOffset=2, PilotToDataPowerOffsetAux=4.685, NrIterations=4,
                 Velocity_kmh=500, PowerDelayProfile="VehicularA", DopplerModel="Jakes", Paths=200,
                 schemes=("aux", "cod", "ofdm"), max_batch=256, device=0, seed=0, verbose=False, setup="device",
                 estimator="auto"):
        if setup not in ("device", "host"):
            raise ValueError("setup must be 'device' (DS.m:208-313 on the GPU) or 'host' (NumPy, cross-check)")
        if estimator not in ("auto", "tiles", "factored", "factored_exact"):
            raise ValueError("estimator must be 'auto', 'tiles', 'factored' or 'factored_exact' (DeviceContext.set_estimator_mode)")
        self.setup_mode = setup
        self.estimator = estimator
        self.setup_times = {}
        self.p = dict(M_SNR_dB=tuple(M_SNR_dB), NrRepetitions=NrRepetitions, ZeroThresholdSparse=ZeroThresholdSparse,
                      L=L, F=F, SamplingRate=SamplingRate, NrSubframes=NrSubframes,
                      QAM_ModulationOrder=QAM_ModulationOrder, PilotToDataPowerOffset=PilotToDataPowerOffset,
                      PilotToDataPowerOffsetAux=PilotToDataPowerOffsetAux, NrIterations=NrIterations,
                      Velocity_kmh=Velocity_kmh, PowerDelayProfile=PowerDelayProfile, DopplerModel=DopplerModel,
                      Paths=Paths)
        self.schemes = tuple(schemes)
        self.seed, self.verbose = seed, verbose
        self.ctx = DeviceContext(device)
        self._setup(max_batch)
        if estimator != "auto":                                 # the factors come from the device-side setup (setup="device")
            self.ctx.set_estimator_mode(estimator)

    @classmethod
    def paper(cls, **kw):
        """DS.m:42-46."""
        base = dict(M_SNR_dB=tuple(range(10, 41, 2)), NrRepetitions=1000, SamplingRate=15e3 * 14 * 14, NrSubframes=2)
        base.update(kw)
        return cls(**base)

    def _log(self, *a):
        if self.verbose:
            print(*a, flush=True)

    # ------------------------------------------------------------------ DS.m:50-313
    def _tick(self, key, t0):
        self.setup_times[key] = self.setup_times.get(key, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def _setup(self, max_batch):
        p, ctx = self.p, self.ctx
        t0 = time.perf_counter()
        L, F, fs, nsf = p["L"], p["F"], p["SamplingRate"], p["NrSubframes"]
        self.FBMC = FBMC(L, 30 * nsf, F, fs, 0, False, "Hermite-OQAM", 8, 0, True)            # DS.m:51-62
        zg = ((self.FBMC.Nr["SamplesTotal"] - (_round_half_away((1 / 15e3 / 14) * fs) + _round_half_away(fs / 15e3))
               * 14 * nsf) / 2) / fs                                                          # DS.m:66
        self.OFDM = OFDM(L, 14 * nsf, F, fs, 0, False, 1 / 15e3 / 14, zg)                     # DS.m:67-76
        if self.OFDM.Nr["SamplesTotal"] != self.FBMC.Nr["SamplesTotal"]:                      # DS.m:79-81
            raise ValueError("Total number of samples must be the same for OFDM and FBMC.")
        N = self.N = self.OFDM.Nr["SamplesTotal"]
        self.PAM = SignalConstellation(int(round(np.sqrt(p["QAM_ModulationOrder"]))), "PAM")  # DS.m:86-87
        self.QAM = SignalConstellation(p["QAM_ModulationOrder"], "QAM")
        pm_o, pm_f, pm_aux = self._pilot_matrices(L, nsf)
        self.PilotMatrix_OFDM, self.PilotMatrix_FBMC, self.AuxilaryPilotMatrix_FBMC = pm_o, pm_f, pm_aux
        pmo, pmf, pma = (m.reshape(-1, order="F") for m in (pm_o, pm_f, pm_aux))
        G_F, G_O = self.FBMC.GetTXMatrix(), self.OFDM.GetTXMatrix()                           # DS.m:191-195
        Q_F, Q_O = self.FBMC.GetRXMatrix().conj().T, self.OFDM.GetRXMatrix().conj().T
        P = self.NrPilotSymbols = int(np.sum(pmo == 1))
        use_fbmc = any(s in self.schemes for s in ("aux", "cod"))
        sch = {}
        if use_fbmc:
            D0 = self.FBMC.GetFBMCMatrix()
        if "aux" in self.schemes:
            self.AuxiliaryMethod = a = _IIC("Auxiliary", pm_aux, D0, 28, p["PilotToDataPowerOffsetAux"])   # DS.m:116-122
            sch["aux"] = dict(wf="F", C=a.PrecodingMatrix, dpr=a.DataPowerReduction, nD=a.NrDataSymbols,
                              kappa=a.PilotToDataPowerOffset * a.DataPowerReduction, const="PAM",
                              detect="select_real", data_pos=np.flatnonzero(pma == 0))
        if "cod" in self.schemes:
            self.CodingMethod = c = _IIC("Coding", pm_f, D0, 20, 2 * p["PilotToDataPowerOffset"])          # DS.m:123-129
            sch["cod"] = dict(wf="F", C=c.PrecodingMatrix, dpr=c.DataPowerReduction, nD=c.NrDataSymbols,
                              kappa=c.PilotToDataPowerOffset * c.DataPowerReduction, const="PAM",
                              detect="despread_real", data_pos=None)
        if "ofdm" in self.schemes:
            Ko, nD = pmo.size, int(np.sum(pmo == 0))
            pmap = np.zeros((Ko, Ko))                                                         # DS.m:134-138
            pmap[np.flatnonzero(pmo == 1), np.arange(P)] = np.sqrt(p["PilotToDataPowerOffset"])
            pmap[np.flatnonzero(pmo == 0), P + np.arange(nD)] = 1.0
            pmap /= np.sqrt(np.mean(np.sum(pmap ** 2, axis=1)))
            dpr = Ko / (P * p["PilotToDataPowerOffset"] + nD)
            self.PilotMapping_OFDM, self.DataPowerReduction_OFDM = pmap, dpr
            sch["ofdm"] = dict(wf="O", C=pmap.astype(np.complex128), dpr=dpr, nD=nD,
                               kappa=p["PilotToDataPowerOffset"] * dpr, const="QAM", detect="select_complex",
                               data_pos=np.flatnonzero(pmo == 0))
        # ---- "NoEdge" masks, DS.m:145-172
        ct_f = np.zeros(pm_f.shape); ct_f[4:-4, 10:-10] = 1
        ct_o = np.zeros(pm_o.shape); ct_o[4:-4, 5:-5] = 1
        ctf, cto = ct_f.reshape(-1, order="F"), ct_o.reshape(-1, order="F")
        nb = {"PAM": self.PAM.BitMapping.shape[1], "QAM": self.QAM.BitMapping.shape[1]}
        for name, s in sch.items():
            Cd = s["C"][:, P:]
            if name == "aux":
                cons = np.sum(np.abs(Cd[(ctf == 1) & (pma == 0), :]), axis=0) > s["dpr"] * 0.9
            elif name == "cod":
                cons = ~np.any(Cd[ctf == 0, :] != 0, axis=0)
            else:
                cons = np.sum(np.abs(Cd[(cto == 1) & (pmo == 0), :]), axis=0) > s["dpr"] * 0.9
            s["considered_bits"] = np.repeat(cons, nb[s["const"]]).astype(np.uint8)
            s["nbits"] = nb[s["const"]]
        self.sch = sch
        t0 = self._tick("modem_and_precoders_host", t0)
        # ---- channel, DS.m:176-186
        fD = p["Velocity_kmh"] / 3.6 * 2.5e9 / 2.998e8
        self.ChannelModel = chan = FastFading(fs, p["PowerDelayProfile"], N, fD, p["DopplerModel"], p["Paths"],
                                              1, 1, False, create_device=False)
        pdp = chan.Implementation["PowerDelayProfileNormalized"]
        self.Pn = np.array([fs / (F * L) * 10 ** (-snr / 10) for snr in p["M_SNR_dB"]])      # DS.m:243,398
        # ---- device context: static operands
        ctx.set_channel(N, pdp, fD, chan.PHY["dt"], p["Paths"], p["DopplerModel"])
        wfs = {}
        if use_fbmc:
            wfs["F"] = dict(G=G_F, Q=Q_F, pil=np.flatnonzero(pmf == 1))
        if "ofdm" in self.schemes:
            wfs["O"] = dict(G=G_O, Q=Q_O, pil=np.flatnonzero(pmo == 1))
        for w, d in wfs.items():
            ctx.set_waveform(w, d["G"], d["Q"])
        if use_fbmc:
            # the polyphase description of the same modem (FBMC.m:61-160): with it the perfect-CSI pass of the loop body applies
            # G and Q' as Modulation / Demodulation (k_perfect_fbmc) instead of as GEMMs; the library checks that it reproduces G, Q
            self.FBMC._set_modem(ctx)
        if "ofdm" in self.schemes:
            self.OFDM._set_modem(ctx)
        ctx.set_constellation("PAM", self.PAM.SymbolMapping, self.PAM.BitMapping)
        ctx.set_constellation("QAM", self.QAM.SymbolMapping, self.QAM.BitMapping)
        ctx.set_snr(self.Pn)
        ctx.finalize(max(P, 1))
        t0 = self._tick("device_static_operands", t0)
        self.wfs = wfs
        self._pdp, self._power_cache, self._schemes_on_device = pdp, {}, False
        self._estimator_setup(max_batch)

    def set_velocity(self, Velocity_kmh):
        """BASELINE.json config 4 (high-mobility sweep): only the time correlation R_t depends on the velocity
        (FF.m:333), so modem matrices, precoders and every static device operand are kept; the correlation and MMSE
        matrices (DS.m:208-313) are rebuilt on the device and the channel's Doppler shift is updated."""
        p = self.p
        p["Velocity_kmh"] = Velocity_kmh
        fD = Velocity_kmh / 3.6 * 2.5e9 / 2.998e8
        self.ChannelModel = FastFading(p["SamplingRate"], p["PowerDelayProfile"], self.N, fD, p["DopplerModel"],
                                       p["Paths"], 1, 1, False, create_device=False)
        self.setup_times = {}
        self._keep_setup_buffers = True
        t0 = time.perf_counter()
        self.ctx.set_channel(self.N, self._pdp, fD, self.ChannelModel.PHY["dt"], p["Paths"], p["DopplerModel"])
        self.ctx.finalize(max(self.NrPilotSymbols, 1))
        self._tick("set_channel_and_finalize", t0)
        self._estimator_setup(self.max_batch_requested)

    def _estimator_setup(self, max_batch):
        """Correlation matrices (DS.m:208-268) and MMSE matrices (DS.m:277-313)."""
        p, ctx, sch, wfs, pdp, P = self.p, self.ctx, self.sch, self.wfs, self._pdp, self.NrPilotSymbols
        self.max_batch_requested = max_batch
        t0 = time.perf_counter()
        thr = 10.0 ** (-p["ZeroThresholdSparse"])
        rt = self.ChannelModel.GetTimeCorrelation()[0]
        taps = np.flatnonzero(pdp)
        for w, d in wfs.items():
            self._log("correlation matrices, waveform", w)
            if self.setup_mode == "device":
                d["R_hP"], d["n_sup"] = ctx.setup_correlations(w, d["pil"], rt, thr)
            else:
                d["R_hP"], d["sup"], d["R_sup"] = self._pilot_correlations(d, pdp, taps, rt, thr, w)
        t0 = self._tick("correlations_%s" % self.setup_mode, t0)
        if not self._schemes_on_device:                       # (the correlation pass needs a finalized context)
            for name, s in sch.items():
                d = wfs[s["wf"]]
                ctx.set_scheme(name, s["wf"], s["C"], d["pil"], s["data_pos"], s["kappa"], s["dpr"], s["detect"],
                               s["const"], s["considered_bits"])
            self._schemes_on_device = True
            t0 = self._tick("device_schemes", t0)
        for name, s in sch.items():
            d = wfs[s["wf"]]
            R_nn = d["R_hP"].copy()
            qn = np.zeros(P)
            for ip in range(P):                                                                # DS.m:222-234
                q = d["Q"][:, d["pil"][ip]]
                R_nn[ip, ip] = self._pilot_power(name, ip, q, s, d, pdp, taps, rt)
                qn[ip] = np.real(np.vdot(q, q))
            t0 = self._tick("pilot_power_host", t0)
            K2 = d["G"].shape[1] ** 2
            for variant in (0, 1):
                Rs = []
                for pn in self.Pn:                                                             # DS.m:238-253,282-285
                    R = R_nn.copy()
                    R[np.arange(P), np.arange(P)] = np.diag(R_nn) + pn * qn / s["kappa"]
                    if variant == 1:
                        R = R - (R_nn - d["R_hP"])
                    Rs.append(np.linalg.pinv(R))
                t0 = self._tick("pinv_host", t0)
                if self.setup_mode == "device":
                    ctx.build_mmse(name, variant, np.stack(Rs), thr)                          # DS.m:283-313 on the GPU
                else:
                    jc, ir, val = [0], [], []
                    for Ri in Rs:
                        Wv = d["R_sup"] @ Ri
                        Wv[np.abs(Wv) < thr] = 0
                        pi, ai = np.nonzero(Wv.T)
                        ir.append(d["sup"][ai].astype(np.int64) + K2 * pi.astype(np.int64))
                        val.append(Wv.T[pi, ai])
                        jc.append(jc[-1] + len(ai))
                    ctx.set_mmse_arrays(name, variant, np.array(jc), np.concatenate(ir), np.concatenate(val))
                t0 = self._tick("mmse_%s" % self.setup_mode, t0)
            s["R_hP_est_noNoise"] = R_nn
        if self.setup_mode == "device" and not getattr(self, "_keep_setup_buffers", False):
            ctx.release_setup()                                # (a velocity sweep keeps R_Dij_hP's buffer: cudaMalloc / cudaFree of 130 MB per step)
        ctx.finalize(max_batch)
        self._tick("finalize", t0)
        self.max_batch = max_batch

    @staticmethod
    def _pilot_matrices(L, nsf):
        """DS.m:91-113."""
        pm_o = np.zeros((L, 14))
        pm_o[1::12, 1::7] = 1; pm_o[4::12, 5::7] = 1; pm_o[7::12, 1::7] = 1; pm_o[10::12, 5::7] = 1
        pm_o = np.tile(pm_o, (1, nsf))
        pm_f = np.zeros((L, 30))
        pm_f[1::12, 2::16] = 1; pm_f[4::12, 10::16] = 1; pm_f[7::12, 3::16] = 1; pm_f[10::12, 11::16] = 1
        pm_f = np.tile(pm_f, (1, nsf))
        aux = pm_f.copy()
        for a, b in zip(*np.nonzero(pm_f)):
            aux[a + 1, b] = aux[a - 1, b] = aux[a, b + 1] = aux[a, b - 1] = -1
        return pm_o, pm_f, aux

    # R_vecH addresses H(:)[a*(N+1)+m] for tap m and column a (FF.m:377); entries past the bottom of
    # a column wrap to the top of the next one and those beyond N^2 are cropped (FF.m:406).
    def _positions(self, m):
        N = self.N
        lin = np.arange(N) * (N + 1) + m
        ok = lin < N * N
        lin = np.where(ok, lin, 0)
        return lin % N, lin // N, ok

    def _pilot_correlations(self, d, pdp, taps, rt, thr, wname):
        """R_hP (DS.m:213), support and values of R_Dij_hP (DS.m:260-267) of one waveform.  The
        P pseudo-channels  M_p = reshape(R_vecH*kron(g_p.',q_p')',N,N)  are banded; their
        Q' M_p G run as one batched K2 launch."""
        N, G, Q, pil = self.N, d["G"], d["Q"], d["pil"]
        P, K, Lt = len(pil), G.shape[1], len(pdp)
        hps = np.zeros((P, N, Lt), dtype=np.complex128)
        corner = []                                   # wrapped (upper-triangular) entries: (p, row, col, value)
        n_idx = np.arange(N)
        for ip in range(P):
            q, g = Q[:, pil[ip]], G[:, pil[ip]]
            for m in taps:
                row, col, ok = self._positions(m)
                zeta = np.where(ok, q[row] * np.conj(g[col]), 0)
                nz = np.flatnonzero(zeta)
                mu = pdp[m] * (rt[(N - 1) + n_idx[:, None] - nz[None, :]] @ zeta[nz])
                reg = ok & (n_idx + m < N)
                hps[ip, row[reg], m] = mu[reg]
                for a in np.flatnonzero(ok & ~reg):
                    corner.append((ip, row[a], col[a], mu[a]))
        self.ctx.set_impulse_response(hps)
        R_hP = np.zeros((P, P), dtype=np.complex128)
        sup_mask = np.zeros(K * K, dtype=bool)
        cols = []
        for ip in range(P):
            Dp, _ = self.ctx.transmission_matrix(wname, ip)
            for (jp, r, c, v) in corner:                                                       # exactness of FF.m:377
                if jp == ip and v != 0:
                    Dp += v * np.outer(np.conj(Q[r, :]), G[c, :])
            R_hP[:, ip] = Dp[pil, pil]
            vec = Dp.reshape(-1, order="F")
            vec[np.abs(vec) < thr] = 0
            sup_mask |= vec != 0
            cols.append(vec)
        sup = np.flatnonzero(sup_mask)
        R_sup = np.stack([v[sup] for v in cols], axis=1)
        return R_hP, sup, R_sup

    def _pilot_power(self, name, ip, q, s, d, pdp, taps, rt):
        """abs(sum(sum((GA.'*(Temp*R_vecH*Temp')).*GA',2))) with Temp = kron(I, q')/sqrt(kappa), GA = G*C, DS.m:203-205,
        224-233: the total power received at one pilot position from all precoded unit-power symbols,
            sum_m pdp_m / kappa * sum_{a,a'} R_t[a-a'] q[r_a] conj(q[r_a']) B[c_a', c_a],   B = GA GA^H.
        B over the pilot's sample window depends on neither tap nor velocity: it is built once per (scheme, pilot) from
        the sparse precoder and cached (a velocity sweep re-uses it)."""
        N = self.N
        key = (name, ip)
        if key not in self._power_cache:
            per_tap = []
            X = []
            for m in taps:
                row, col, ok = self._positions(m)
                sel = np.flatnonzero(ok & (q[row] != 0))
                per_tap.append((m, sel, row[sel], col[sel]))
                X.append(col[sel])
            rows = np.unique(np.concatenate(X))
            if "GA_sparse" not in s:
                s["GA_sparse"] = sp.csr_matrix(sp.csr_matrix(d["G"]) @ sp.csc_matrix(s["C"]))   # DS.m:203-205
            GAx = s["GA_sparse"][rows, :]
            GAx = GAx[:, np.flatnonzero(np.diff(GAx.tocsc().indptr))].toarray()
            B = GAx @ GAx.conj().T
            pos = np.full(N, -1); pos[rows] = np.arange(len(rows))
            # only R_t depends on the velocity, and it enters through the lag a - a' alone: fold everything else into one
            # weight per lag, w[lag] = sum_m pdp_m / kappa sum_{a - a' = lag} q[r_a] conj(q[r_a']) B[c_a', c_a]
            w = np.zeros(2 * N - 1, dtype=np.complex128)
            for m, sel, r, c in per_tap:
                qq = q[r]
                Bc = B[np.ix_(pos[c], pos[c])].T                               # B[c_a', c_a] indexed [a, a']
                Mm = (pdp[m] / s["kappa"]) * (qq[:, None] * np.conj(qq)[None, :]) * Bc
                lag = ((N - 1) + sel[:, None] - sel[None, :]).ravel()
                w += np.bincount(lag, weights=Mm.real.ravel(), minlength=2 * N - 1) + 1j * np.bincount(lag, weights=Mm.imag.ravel(), minlength=2 * N - 1)
            nz = np.flatnonzero(w)
            self._power_cache[key] = (nz, w[nz])
        nz, w = self._power_cache[key]
        return abs(np.dot(rt[nz], w))

    # ------------------------------------------------------------------ DS.m:350-565
    def run(self, NrRepetitions=None, NrIterations=None, seed=None, draws=None, first_rep=0):
        """Monte-Carlo loop.  Returns (ber, err): the 24 BER arrays of DS.m:322-345 keyed by the
        reference's variable names (S x reps [x I]) and the raw error counts
        err[rep, snr, it, scheme, csi, edge]."""
        R = self.p["NrRepetitions"] if NrRepetitions is None else NrRepetitions
        I = self.p["NrIterations"] if NrIterations is None else NrIterations
        seed = self.seed if seed is None else seed
        err = np.zeros((R, len(self.Pn), I + 1, 3, 2, 2), dtype=np.uint32)
        for r0 in range(0, R, self.max_batch):
            n = min(self.max_batch, R - r0)
            if draws is not None:
                st, keep = self.ctx.pack_draws(draws[r0:r0 + n])
                err[r0:r0 + n] = self.ctx.run_batch(n, I, st)
            else:
                err[r0:r0 + n] = self.ctx.run_batch(n, I, None, seed=seed, first_rep=first_rep + r0)
        return self.ber_arrays(err), err

    def run_totals(self, NrRepetitions, NrIterations=None, seed=None, first_rep=0):
        """The loop for `NrRepetitions` seeded realizations with the per-realization counters left on the device and
        summed there (chest_multi_run over this simulation's one context): returns the 64-bit totals
        [snr, it, scheme, csi, edge] -- what a sharded run reduces over the GPUs."""
        from .context import MultiDevice
        I = self.p["NrIterations"] if NrIterations is None else NrIterations
        if getattr(self, "_multi", None) is None:
            self._multi = MultiDevice([self.ctx])
        if NrRepetitions <= 0:
            return np.zeros((len(self.Pn), I + 1, 3, 2, 2), dtype=np.uint64)
        _, tot, _ = self._multi.run(NrRepetitions, I, seed=self.seed if seed is None else seed, first_rep=first_rep, want_err=False)
        return tot

    def ber_arrays(self, err):
        nb = self.ctx.bit_counts()
        out = {}
        for name in self.sch:
            sid = SCHEME_ID[name]
            for ci, ctag in ((0, ""), (1, "_PerfectCSI")):
                for ei, etag in ((0, ""), (1, "_NoEdge")):
                    e = np.transpose(err[:, :, :, sid, ci, ei], (1, 0, 2)) / float(nb[sid, ei])   # S x reps x (1+I)
                    out["BER_%s_OneTapEqualizer%s%s" % (BER_NAMES[name], ctag, etag)] = e[:, :, 0]
                    key = ("BER_%s_InterferenceCancellation%s" if ci == 0 else
                           "BER_%s_PerfectCSI_InterferenceCancellation%s") % (BER_NAMES[name], etag)
                    out[key] = e[:, :, 1:]
        return out

    def close(self):
        if getattr(self, "_multi", None) is not None:
            self._multi.close()
            self._multi = None
        self.ctx.close()


class SimpleVersionSimulation:
    """SimpleVersion_DoublyFlat.m with its double loop (SV.m:89-176) replaced by batched device launches
    (chest_sv_run_batch): FFT-form FBMC / OFDM modem, doubly-flat Rayleigh channel, LS pilot estimates, interpolation,
    one-tap equalisation, BER.  The parameter block is SV.m:12-82."""

    def __init__(self, M_SNR_OFDM_dB=tuple(range(0, 31, 5)), NrRepetitions=1000, QAM_ModulationOrder=16,
                 NrSubcarriers=12, max_batch=4096, device=0, seed=0):
        from .estimation import PilotSymbolAidedChannelEstimation as PSACE
        self.M_SNR_OFDM_dB, self.NrRepetitions, self.seed = tuple(M_SNR_OFDM_dB), NrRepetitions, seed
        L = NrSubcarriers
        fs = 15e3 * 14 * 12
        self.FBMC = FBMC(L, 30, 15e3, fs, 15e3 * 20, False, "Hermite-OQAM", 8, 0, True)                       # SV.m:17-28
        self.OFDM = OFDM(L, 15, 15e3, fs, 15e3 * 20, False, 0, (8 - 1 / 2) * 1 / 15e3 * 1 / 2)                # SV.m:31-40
        self.PAM = SignalConstellation(int(round(np.sqrt(QAM_ModulationOrder))), "PAM")                       # SV.m:43-44
        self.QAM = SignalConstellation(QAM_ModulationOrder, "QAM")
        self.ChannelEstimation_FBMC = ce_f = PSACE("Diamond", [[L, 6], [30, 8]], "linear")                    # SV.m:47-56
        self.ChannelEstimation_OFDM = ce_o = PSACE("Diamond", [[L, 6], [15, 4]], "linear")                    # SV.m:57-66
        D0 = self.FBMC.GetFBMCMatrix()
        self.AuxiliaryMethod = aux = _IIC("Auxiliary", ce_f.GetAuxiliaryMatrix(1), D0, 16, 2)                 # SV.m:71-75
        self.CodingMethod = cod = _IIC("Coding", ce_f.PilotMatrix, D0, 16, 2)                                 # SV.m:76-80
        pmf = ce_f.PilotMatrix.reshape(-1, order="F"); pmo = ce_o.PilotMatrix.reshape(-1, order="F")
        pma = aux.PilotMatrix.reshape(-1, order="F")
        Pf, Po, Ko = ce_f.NrPilotSymbols, ce_o.NrPilotSymbols, pmo.size
        map_o = np.zeros((Ko, Ko))                                                                             # SV.m:113-115
        map_o[np.flatnonzero(pmo == 1), np.arange(Po)] = 1.0
        map_o[np.flatnonzero(pmo == 0), Po + np.arange(Ko - Po)] = 1.0
        self.interp_f, self.interp_o = ce_f.GetInterpolationMatrix(), ce_o.GetInterpolationMatrix()            # PSACE.m:171-184
        ctx = self.ctx = DeviceContext(device)
        self.FBMC._set_modem(ctx); self.OFDM._set_modem(ctx)
        ctx.set_constellation("PAM", self.PAM.SymbolMapping, self.PAM.BitMapping)
        ctx.set_constellation("QAM", self.QAM.SymbolMapping, self.QAM.BitMapping)
        nbp, nbq = self.PAM.BitMapping.shape[1], self.QAM.BitMapping.shape[1]
        ctx.K["F"], ctx.K["O"] = pmf.size, Ko
        ctx.set_scheme("aux", "F", aux.PrecodingMatrix, np.flatnonzero(pmf == 1), np.flatnonzero(pma == 0),
                       aux.PilotToDataPowerOffset * aux.DataPowerReduction, aux.DataPowerReduction, "select_real", "PAM",
                       np.ones(aux.NrDataSymbols * nbp, dtype=np.uint8))                                        # SV.m:138,148
        ctx.set_scheme("cod", "F", cod.PrecodingMatrix, np.flatnonzero(pmf == 1), None, cod.PilotToDataPowerOffset, 1.0,
                       "despread_real", "PAM", np.ones(cod.NrDataSymbols * nbp, dtype=np.uint8))                # SV.m:139,149
        ctx.set_scheme("ofdm", "O", map_o.astype(np.complex128), np.flatnonzero(pmo == 1), np.flatnonzero(pmo == 0), 1.0, 1.0,
                       "select_complex", "QAM", np.ones((Ko - Po) * nbq, dtype=np.uint8))                       # SV.m:140,152
        ctx.set_interpolation("aux", self.interp_f); ctx.set_interpolation("cod", self.interp_f)
        ctx.set_interpolation("ofdm", self.interp_o)
        ctx.finalize(max_batch)
        self.max_batch = max_batch
        self.n_bits = np.array([aux.NrDataSymbols * nbp, cod.NrDataSymbols * nbp, cod.NrDataSymbols * nbp,
                                (Ko - Po) * nbq, (Ko - Po) * nbq], dtype=np.float64)

    def noise_power(self, snr_db):
        o = self.OFDM
        return o.PHY["SamplingRate"] / (o.PHY["SubcarrierSpacing"] * o.Nr["Subcarriers"]) * 10 ** (-np.asarray(snr_db, dtype=float) / 10)   # SV.m:92

    def run(self, NrRepetitions=None, seed=None):
        """Returns the five BER arrays of SV.m:84-88 (n_SNR x NrRepetitions each) keyed by the script's variable names."""
        R = self.NrRepetitions if NrRepetitions is None else NrRepetitions
        nS = len(self.M_SNR_OFDM_dB)
        pn = np.tile(self.noise_power(self.M_SNR_OFDM_dB), R)                 # body = rep * nS + snr (the script's loop order)
        err = np.zeros((R * nS, 5), dtype=np.uint32)
        for b0 in range(0, R * nS, self.max_batch):
            n = min(self.max_batch, R * nS - b0)
            err[b0:b0 + n] = self.ctx.sv_run_batch(pn[b0:b0 + n], None, seed=self.seed if seed is None else seed, first_body=b0)
        ber = err.reshape(R, nS, 5).transpose(1, 0, 2) / self.n_bits[None, None, :]
        names = ("BER_FBMC_Aux", "BER_FBMC_Cod", "BER_FBMC_perfect", "BER_OFDM", "BER_OFDM_perfect")
        return {n: ber[:, :, k] for k, n in enumerate(names)}, err

    def close(self):
        self.ctx.close()
