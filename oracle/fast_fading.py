"""Oracle restatement of +Channel/FastFading.m: constructor, NewRealization (Jakes / Uniform /
Discrete-* / block fading / AWGN), GetConvolutionMatrix, Convolution, GetTimeCorrelation,
GetCorrelationMatrix.  1x1 antennas only (what the scripts use).
Test infrastructure only -- see oracle/__init__.py."""
import numpy as np
import scipy.sparse as sp
from scipy.special import j0

# FastFading.m:56-107 -- [relative power dB; relative delay s]
_PDP_TABLES = {
    "Flat": ([0.0], [0.0]),
    "AWGN": ([0.0], [0.0]),
    "PedestrianA": ([0, -9.7, -19.2, -22.8], [0, 110e-9, 190e-9, 410e-9]),
    "PedestrianB": ([0, -0.9, -4.9, -8, -7.8, -23.9], [0, 200e-9, 800e-9, 1200e-9, 2300e-9, 3700e-9]),
    "VehicularA": ([0, -1, -9, -10, -15, -20], [0, 310e-9, 710e-9, 1090e-9, 1730e-9, 2510e-9]),
    "VehicularB": ([-2.5, 0, -12.8, -10, -25.2, -16], [0, 300e-9, 8900e-9, 12900e-9, 17100e-9, 20000e-9]),
    "ExtendedPedestrianA": ([0, -1, -2, -3, -8, -17.2, -20.8],
                            [0, 30e-9, 70e-9, 90e-9, 110e-9, 190e-9, 410e-9]),
    "ExtendedVehicularA": ([0, -1.5, -1.4, -3.6, -0.6, -9.1, -7, -12, -16.9],
                           [0, 30e-9, 150e-9, 310e-9, 370e-9, 710e-9, 1090e-9, 1730e-9, 2510e-9]),
}
_TDL = {
    "TDL-A": ([-13.4, 0, -2.2, -4, -6, -8.2, -9.9, -10.5, -7.5, -15.9, -6.6, -16.7, -12.4, -15.2, -10.8,
               -11.3, -12.7, -16.2, -18.3, -18.9, -16.6, -19.9, -29.7],
              [0.0000, 0.3819, 0.4025, 0.5868, 0.4610, 0.5375, 0.6708, 0.5750, 0.7618, 1.5375, 1.8978,
               2.2242, 2.1718, 2.4942, 2.5119, 3.0582, 4.0810, 4.4579, 4.5695, 4.7966, 5.0066, 5.3043,
               9.6586]),
    "TDL-B": ([0, -2.2, -4, -3.2, -9.8, -1.2, -3.4, -5.2, -7.6, -3, -8.9, -9, -4.8, -5.7, -7.5, -1.9, -7.6,
               -12.2, -9.8, -11.4, -14.9, -9.2, -11.3],
              [0.0000, 0.1072, 0.2155, 0.2095, 0.2870, 0.2986, 0.3752, 0.5055, 0.3681, 0.3697, 0.5700,
               0.5283, 1.1021, 1.2756, 1.5474, 1.7842, 2.0169, 2.8294, 3.0219, 3.6187, 4.1067, 4.2790,
               4.7834]),
    "TDL-C": ([-4.4, -1.2, -3.5, -5.2, -2.5, 0, -2.2, -3.9, -7.4, -7.1, -10.7, -11.1, -5.1, -6.8, -8.7,
               -13.2, -13.9, -13.9, -15.8, -17.1, -16, -15.7, -21.6, -22.8],
              [0, 0.2099, 0.2219, 0.2329, 0.2176, 0.6366, 0.6448, 0.6560, 0.6584, 0.7935, 0.8213, 0.9336,
               1.2285, 1.3083, 2.1704, 2.7105, 4.2589, 4.6003, 5.4902, 5.6077, 6.3065, 6.6374, 7.0427,
               8.6523]),
}


def _mround(x):
    return np.floor(np.abs(x) + 0.5) * np.sign(x)


def resample_power_delay_profile(PowerDelayProfile, SamplingRate):
    """FastFading.m:47-131.  Returns (PowerDelayProfile over taps, normalized PDP, IndexDelayTaps 0-based)."""
    dt = 1.0 / SamplingRate
    if isinstance(PowerDelayProfile, str):
        name = PowerDelayProfile
        if name[:3] == "TDL":                                                   # :49-54
            p1, p2 = name.find("_"), name.find("ns")
            spread = float(name[p1 + 1:p2]) * 1e-9
            p_db, rel = _TDL[name[:5]]
            delays = spread * np.asarray(rel)
        else:
            if name not in _PDP_TABLES:
                raise ValueError("Power delay profile model not supported!")
            p_db, delays = _PDP_TABLES[name]
            delays = np.asarray(delays, dtype=float)
        idx = (_mround(delays / dt)).astype(int)                                # :111 (0-based)
        pdp = np.zeros(idx.max() + 1)
        for i, d in enumerate(idx):                                             # :117-121
            pdp[d] += 10.0 ** (p_db[i] / 10.0)
    else:
        pdp = np.asarray(PowerDelayProfile, dtype=float).reshape(-1)            # :124
    pdp_norm = pdp / pdp.sum()                                                  # :129
    taps = np.flatnonzero(pdp)                                                  # :131
    return pdp, pdp_norm, taps


class FastFading:
    """FastFading.m:25-192."""

    def __init__(self, SamplingRate, PowerDelayProfile, SamplesTotal, MaximumDopplerShift,
                 DopplerModel, Paths, nTxAntennas=1, nRxAntennas=1,
                 WarningIfSampleRateDoesNotMatch=False, rng=None):
        if nTxAntennas != 1 or nRxAntennas != 1:
            raise NotImplementedError("oracle covers 1x1 antennas")
        self.PHY = {"SamplingRate": float(SamplingRate), "MaximumDopplerShift": float(MaximumDopplerShift),
                    "dt": 1.0 / SamplingRate}
        self.Nr = {"SamplesTotal": int(SamplesTotal), "txAntennas": 1, "rxAntennas": 1}
        self.Implementation = {"PowerDelayProfile": PowerDelayProfile}
        self._rng = rng if rng is not None else np.random.default_rng(0)
        if isinstance(PowerDelayProfile, str) and PowerDelayProfile == "AWGN" and MaximumDopplerShift > 0:
            self.PHY["MaximumDopplerShift"] = 0.0                               # :65-68
        pdp, pdp_norm, taps = resample_power_delay_profile(PowerDelayProfile, SamplingRate)
        self.PHY["PowerDelayProfile"] = pdp
        self.Implementation["PowerDelayProfileNormalized"] = pdp_norm
        self.Implementation["IndexDelayTaps"] = taps
        N, fD = self.Nr["SamplesTotal"], self.PHY["MaximumDopplerShift"]
        discrete = DopplerModel[:8] == "Discrete"
        if fD / (SamplingRate / N) <= 0.5 and discrete and fD > 0:              # :146-149
            self.PHY["MaximumDopplerShift"] = fD = 0.0
        if fD > 0:                                                              # :151-182
            self.PHY["DopplerModel"] = DopplerModel
            self.Implementation["UseDiscreteDopplerSpectrum"] = discrete
            if discrete:
                df = SamplingRate / N
                n_shift = int(np.ceil(fD / df))
                pts = df * (np.arange(-n_shift - 1, n_shift + 1) + 0.5)
                pts = np.clip(pts, -fD, fD)
                if DopplerModel == "Discrete-Jakes":
                    spec = np.arcsin(pts[1:] / fD) - np.arcsin(pts[:-1] / fD)
                elif DopplerModel == "Discrete-Uniform":
                    spec = pts[1:] - pts[:-1]
                else:
                    raise ValueError("Doppler spectrum not supported")
                spec = spec / spec.sum()
                self.Implementation["DiscreteDopplerSpectrum"] = np.repeat(spec[:, None], len(taps), axis=1)
            else:
                self.Nr["Paths"] = int(Paths)
        self.ImpulseResponse = None
        self.NewRealization()                                                   # :184

    # ---------------------------------------------------------------- realization
    def NewRealization(self, doppler_u=None, phase_u=None, gauss=None):
        """FastFading.m:194-250.  Optional explicit draws replace MATLAB's RNG:
        doppler_u / phase_u: (n_taps, Paths) uniforms for the Jakes/Uniform branch (the reference
        draws rand([T 1 Paths]) twice, Doppler first, FastFading.m:227,233);
        gauss: complex standard normals for the Discrete-* (2*NrShifts+1, n_taps) or the
        block-fading (Lt,) branch."""
        Imp, N = self.Implementation, self.Nr["SamplesTotal"]
        pdp_norm, taps = Imp["PowerDelayProfileNormalized"], Imp["IndexDelayTaps"]
        fD, dt = self.PHY["MaximumDopplerShift"], self.PHY["dt"]
        if isinstance(Imp["PowerDelayProfile"], str) and Imp["PowerDelayProfile"] == "AWGN":
            self.ImpulseResponse = np.ones((1, 1), dtype=complex)               # :197-198
            return
        if fD > 0:
            h = np.zeros((N, len(pdp_norm)), dtype=complex)                     # :201
            if Imp["UseDiscreteDopplerSpectrum"]:                               # :203-221
                spec = Imp["DiscreteDopplerSpectrum"]
                ns = (spec.shape[0] - 1) // 2
                if gauss is None:
                    gauss = (self._rng.standard_normal((2 * ns + 1, len(taps)))
                             + 1j * self._rng.standard_normal((2 * ns + 1, len(taps))))
                g = N / np.sqrt(2) * np.asarray(gauss) * np.sqrt(pdp_norm[taps])[None, :]
                g1, g2 = g[:ns + 1, :], g[ns + 1:, :]
                spec_in = np.vstack([np.sqrt(spec[ns:, :]) * g1,
                                     np.zeros((N - 2 * ns - 1, len(taps))),
                                     np.sqrt(spec[:ns, :]) * g2])
                h[:, taps] = np.fft.ifft(spec_in, axis=0)
            else:                                                               # :223-239
                P = self.Nr["Paths"]
                if doppler_u is None:
                    doppler_u = self._rng.random((len(taps), P))
                if phase_u is None:
                    phase_u = self._rng.random((len(taps), P))
                if self.PHY["DopplerModel"] == "Jakes":
                    shifts = np.cos(np.asarray(doppler_u) * 2 * np.pi) * fD     # :227
                elif self.PHY["DopplerModel"] == "Uniform":
                    shifts = 2 * (np.asarray(doppler_u) - 0.5) * fD             # :229
                else:
                    raise ValueError("Doppler spectrum not supported")
                t = np.arange(N) * dt                                           # :234
                arg = np.asarray(phase_u)[:, :, None] + shifts[:, :, None] * t[None, None, :]
                tmp = np.exp(1j * 2 * np.pi * arg).sum(axis=1) / np.sqrt(P)     # :235
                h[:, taps] = (np.sqrt(pdp_norm[taps])[:, None] * tmp).T         # :237
            self.ImpulseResponse = h
        else:                                                                   # :241-248
            if gauss is None:
                gauss = self._rng.standard_normal(len(pdp_norm)) + 1j * self._rng.standard_normal(len(pdp_norm))
            self.ImpulseResponse = (1 / np.sqrt(2) * np.sqrt(pdp_norm) * np.asarray(gauss)).reshape(1, -1)

    # ---------------------------------------------------------------- operator
    def GetConvolutionMatrix(self):
        """FastFading.m:276-295: sparse N x N, H[r, r-m] = h[r, m] (r >= m).  Returns the
        {1,1} cell entry as scipy CSC."""
        N = self.Nr["SamplesTotal"]
        taps = self.Implementation["IndexDelayTaps"]
        h = self.ImpulseResponse
        rows, cols, vals = [], [], []
        for m in taps:
            r = np.arange(m, N)
            rows.append(r)
            cols.append(r - m)
            vals.append(h[r, m] if h.shape[0] > 1 else np.full(len(r), h[0, m]))
        return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                             shape=(N, N))

    def Convolution(self, signal):
        """FastFading.m:253-274."""
        signal = np.asarray(signal).reshape(-1)
        n = len(signal)
        if self.PHY["MaximumDopplerShift"] > 0:
            return self.GetConvolutionMatrix()[:n, :n] @ signal                 # :262
        return np.convolve(signal, self.ImpulseResponse[0, :])[:n]              # :269-272

    # ------------------------------------------------------------- statistics
    def GetTimeCorrelation(self):
        """FastFading.m:321-340.  Returns (TimeCorrelation, Time), both length 2N-1."""
        N, dt, fD = self.Nr["SamplesTotal"], self.PHY["dt"], self.PHY["MaximumDopplerShift"]
        time = (np.arange(2 * N - 1) - (N - 1)) * dt
        if fD > 0:
            if self.PHY["DopplerModel"] in ("Jakes", "Discrete-Jakes"):
                return j0(np.pi * 2 * fD * time), time                          # :333
            return np.sinc(2 * fD * time), time                                 # :335
        return np.ones(2 * N - 1), time

    def GetCorrelationMatrix(self):
        """FastFading.m:366-407: R_vecH = E{H(:)H(:)'} as an N^2 x N^2 sparse matrix, built with
        the reference's own linear-index arithmetic (including the wrap of the last Lt-1 columns
        into upper-triangular positions, FastFading.m:377,406)."""
        N = self.Nr["SamplesTotal"]
        pdp_norm = self.Implementation["PowerDelayProfileNormalized"]
        Lt = self.ImpulseResponse.shape[1]                                      # size(ImpulseResponse,2)
        rt, _ = self.GetTimeCorrelation()
        a = np.arange(N)
        toep = rt[(N - 1) + a[:, None] - a[None, :]]                            # :372-373
        icc = (a * (N + 1))[:, None] + np.arange(Lt)[None, :]                   # :377 (0-based linear)
        rows, cols, vals = [], [], []
        for m in range(Lt):
            if pdp_norm[m] == 0:
                continue
            rr = np.repeat(icc[:, m][:, None], N, axis=1)                       # row index: entry (a,m)
            cc = np.repeat(icc[:, m][None, :], N, axis=0)                       # col index: entry (b,m)
            rows.append(rr.reshape(-1))
            cols.append(cc.reshape(-1))
            vals.append((pdp_norm[m] * toep).reshape(-1))                       # :374-376
        rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
        keep = (rows < N * N) & (cols < N * N)                                  # :406
        return sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(N * N, N * N))
