"""Host mirror of the reference's +Channel package: `Channel.FastFading` (FF.m).

The constructor (power-delay-profile tables, resampling to the sampling rate) is host-side
table work, as in the reference; `NewRealization`, `Convolution`, `GetConvolutionMatrix` run on the
GPU through the C ABI: 'Jakes' and 'Uniform' sum-of-sinusoids synthesis (the branch both reference scripts use,
FF.m:223-239), the 'Discrete-Jakes' / 'Discrete-Uniform' IFFT synthesis (FF.m:151-182, 203-221: a pruned inverse DFT
over the 2 n_shift + 1 Doppler bins), block fading (f_D = 0, FF.m:241-248) and 'AWGN' (FF.m:197-198), whose few
draws are made on the host and applied by the same device operator.  nTx x nRx antennas (FF.m:205-206, 223-224,
258-262, 279-285) are a batch of nRx * nTx independent realizations on the device (index rx + nRx * tx)."""
import numpy as np
from scipy.special import j0

from .context import DeviceContext

# relative power (dB) / delay (s) tables, FF.m:56-92
_PDP = {
    "Flat": ((0.0,), (0.0,)),
    "AWGN": ((0.0,), (0.0,)),
    "PedestrianA": ((0, -9.7, -19.2, -22.8), (0, 110e-9, 190e-9, 410e-9)),
    "PedestrianB": ((0, -0.9, -4.9, -8, -7.8, -23.9), (0, 200e-9, 800e-9, 1200e-9, 2300e-9, 3700e-9)),
    "VehicularA": ((0, -1, -9, -10, -15, -20), (0, 310e-9, 710e-9, 1090e-9, 1730e-9, 2510e-9)),
    "VehicularB": ((-2.5, 0, -12.8, -10, -25.2, -16), (0, 300e-9, 8900e-9, 12900e-9, 17100e-9, 20000e-9)),
    "ExtendedPedestrianA": ((0, -1, -2, -3, -8, -17.2, -20.8), (0, 30e-9, 70e-9, 90e-9, 110e-9, 190e-9, 410e-9)),
    "ExtendedVehicularA": ((0, -1.5, -1.4, -3.6, -0.6, -9.1, -7, -12, -16.9),
                           (0, 30e-9, 150e-9, 310e-9, 370e-9, 710e-9, 1090e-9, 1730e-9, 2510e-9)),
}
# 3GPP 38.900 tapped delay lines, FF.m:93-107: (power dB, delay / rms delay spread)
_TDL = {
    "TDL-A": ((-13.4, 0, -2.2, -4, -6, -8.2, -9.9, -10.5, -7.5, -15.9, -6.6, -16.7, -12.4, -15.2, -10.8, -11.3,
               -12.7, -16.2, -18.3, -18.9, -16.6, -19.9, -29.7),
              (0.0000, 0.3819, 0.4025, 0.5868, 0.4610, 0.5375, 0.6708, 0.5750, 0.7618, 1.5375, 1.8978, 2.2242,
               2.1718, 2.4942, 2.5119, 3.0582, 4.0810, 4.4579, 4.5695, 4.7966, 5.0066, 5.3043, 9.6586)),
    "TDL-B": ((0, -2.2, -4, -3.2, -9.8, -1.2, -3.4, -5.2, -7.6, -3, -8.9, -9, -4.8, -5.7, -7.5, -1.9, -7.6, -12.2,
               -9.8, -11.4, -14.9, -9.2, -11.3),
              (0.0000, 0.1072, 0.2155, 0.2095, 0.2870, 0.2986, 0.3752, 0.5055, 0.3681, 0.3697, 0.5700, 0.5283,
               1.1021, 1.2756, 1.5474, 1.7842, 2.0169, 2.8294, 3.0219, 3.6187, 4.1067, 4.2790, 4.7834)),
    "TDL-C": ((-4.4, -1.2, -3.5, -5.2, -2.5, 0, -2.2, -3.9, -7.4, -7.1, -10.7, -11.1, -5.1, -6.8, -8.7, -13.2,
               -13.9, -13.9, -15.8, -17.1, -16, -15.7, -21.6, -22.8),
              (0, 0.2099, 0.2219, 0.2329, 0.2176, 0.6366, 0.6448, 0.6560, 0.6584, 0.7935, 0.8213, 0.9336, 1.2285,
               1.3083, 2.1704, 2.7105, 4.2589, 4.6003, 5.4902, 5.6077, 6.3065, 6.6374, 7.0427, 8.6523)),
}


class FastFading:
    """Channel.FastFading(SamplingRate, PowerDelayProfile, SamplesTotal, MaximumDopplerShift,
    DopplerModel, Paths, nTxAntennas, nRxAntennas, WarningIfSampleRateDoesNotMatch)  (FF.m:25-35)."""

    def __init__(self, SamplingRate, PowerDelayProfile, SamplesTotal, MaximumDopplerShift, DopplerModel,
                 Paths, nTxAntennas=1, nRxAntennas=1, WarningIfSampleRateDoesNotMatch=False, seed=0,
                 create_device=True):
        if DopplerModel not in ("Jakes", "Uniform", "Discrete-Jakes", "Discrete-Uniform") and MaximumDopplerShift > 0:
            raise ValueError("Doppler spectrum not supported")                           # FF.m:230,171
        self.PHY = {"SamplingRate": float(SamplingRate), "MaximumDopplerShift": float(MaximumDopplerShift),
                    "dt": 1.0 / SamplingRate, "DopplerModel": DopplerModel}
        self.Nr = {"SamplesTotal": int(SamplesTotal), "txAntennas": int(nTxAntennas), "rxAntennas": int(nRxAntennas),
                   "Paths": int(Paths)}
        discrete = str(DopplerModel).startswith("Discrete")
        self.Implementation = {"PowerDelayProfile": PowerDelayProfile, "UseDiscreteDopplerSpectrum": False}
        dt = self.PHY["dt"]
        if isinstance(PowerDelayProfile, str):
            name = PowerDelayProfile
            if name.startswith("TDL"):                                                   # FF.m:49-54
                spread = float(name[name.index("_") + 1:name.index("ns")]) * 1e-9
                p_db, rel = _TDL[name[:5]]
                delays = spread * np.asarray(rel)
            elif name in _PDP:
                p_db, delays = _PDP[name]
                delays = np.asarray(delays, dtype=float)
            else:
                raise ValueError("Power delay profile model not supported!")
            idx = np.floor(delays / dt + 0.5).astype(int)                                # round(), FF.m:111
            if WarningIfSampleRateDoesNotMatch and np.sum(np.abs(np.remainder(delays, dt))) > 0:
                import sys; print("Sampling rate does not match the predefined delays of the channel model!", file=sys.stderr)
            pdp = np.bincount(idx, weights=10.0 ** (np.asarray(p_db, dtype=float) / 10.0))   # FF.m:117-121
            self.PHY["DesiredPowerDelayProfiledB"] = np.vstack([p_db, delays])
        else:
            pdp = np.asarray(PowerDelayProfile, dtype=float).reshape(-1)                  # FF.m:124
        self.PHY["PowerDelayProfile"] = pdp
        self.Implementation["PowerDelayProfileNormalized"] = pdp / pdp.sum()              # FF.m:129
        self.Implementation["IndexDelayTaps"] = np.flatnonzero(pdp)                       # FF.m:131
        fD, N = self.PHY["MaximumDopplerShift"], self.Nr["SamplesTotal"]
        if discrete and fD > 0 and fD / (SamplingRate / N) <= 0.5:                        # FF.m:146-149
            import sys; print('Discrete Doppler spectrum: The velocity is so low, that it is set to zero.', file=sys.stderr)
            self.PHY["MaximumDopplerShift"] = fD = 0.0
        if discrete and fD > 0:                                                           # FF.m:153-175
            self.Implementation["UseDiscreteDopplerSpectrum"] = True
            df = SamplingRate / N
            ns = int(np.ceil(fD / df))
            pts = np.clip(df * (np.arange(-ns - 1, ns + 1) + 0.5), -fD, fD)
            spec = (np.arcsin(pts[1:] / fD) - np.arcsin(pts[:-1] / fD)) if DopplerModel == "Discrete-Jakes" else (pts[1:] - pts[:-1])
            spec = spec / spec.sum()
            self.Implementation["DiscreteDopplerSpectrum"] = np.repeat(spec[:, None], len(self.Implementation["IndexDelayTaps"]), axis=1)
        self._seed, self._count = int(seed), 0
        self._ctx = None
        self.ImpulseResponse = None
        if create_device:
            self.NewRealization()                                                         # FF.m:184

    # ------------------------------------------------------------------ device
    def _n_links(self):
        return self.Nr["txAntennas"] * self.Nr["rxAntennas"]

    def _device(self):
        if self._ctx is None:
            self._ctx = DeviceContext()
            # a time-invariant channel (f_D = 0, AWGN) is uploaded as an impulse response; the synthesis parameters are
            # then unused
            model = self.PHY["DopplerModel"] if self.PHY["MaximumDopplerShift"] > 0 else "Jakes"
            self._ctx.set_channel(self.Nr["SamplesTotal"], self.Implementation["PowerDelayProfileNormalized"],
                                  self.PHY["MaximumDopplerShift"], self.PHY["dt"], self.Nr["Paths"], model)
            self._ctx.finalize(self._n_links())
        return self._ctx

    def _as_cells(self, per_link):
        """list over links (rx + nRx * tx) -> array [..., nRx, nTx] like obj.ImpulseResponse(:,:,iRx,iTx)."""
        nT, nR = self.Nr["txAntennas"], self.Nr["rxAntennas"]
        a = np.stack(per_link, axis=-1)
        return a.reshape(a.shape[:-1] + (nT, nR)).swapaxes(-1, -2)

    def NewRealization(self, doppler_u=None, phase_u=None, gauss=None):
        """FF.m:194-250.  Without arguments the draws come from the device generator keyed by (seed, call count);
        explicit draws reproduce exported ones: (T x Paths) uniforms per link for 'Jakes'/'Uniform', (2 n_shift + 1, T)
        complex normals per link for 'Discrete-*', Lt complex normals per link for block fading.  With several antennas the
        leading axis of the explicit draws runs over the links in the reference's loop order (tx outer, rx inner)."""
        ctx = self._device()
        N, pdp = self.Nr["SamplesTotal"], self.Implementation["PowerDelayProfileNormalized"]
        nl = self._n_links()
        siso = nl == 1
        is_awgn = isinstance(self.Implementation["PowerDelayProfile"], str) and self.Implementation["PowerDelayProfile"] == "AWGN"
        if is_awgn:
            h = np.ones((nl, 1, 1), dtype=complex)                                        # FF.m:197-198
            ctx.set_impulse_response(np.broadcast_to(h, (nl, N, len(pdp))).copy())
            self._count += 1
            self.ImpulseResponse = h[0] if siso else self._as_cells(list(h))
            return
        if not self.PHY["MaximumDopplerShift"] > 0:
            # block fading, FF.m:241-248: one complex normal per tap and link, constant over the block
            if gauss is None:
                rng = np.random.default_rng([self._seed, self._count])
                gauss = rng.standard_normal((nl, len(pdp))) + 1j * rng.standard_normal((nl, len(pdp)))
            g = np.asarray(gauss).reshape(nl, len(pdp))
            h = (np.sqrt(pdp) / np.sqrt(2))[None, :] * g
            ctx.set_impulse_response(np.broadcast_to(h[:, None, :], (nl, N, len(pdp))).copy())   # H applied on the GPU as for f_D > 0
            self._count += 1
            self.ImpulseResponse = h[0][None, :] if siso else self._as_cells([x[None, :] for x in h])
            return
        if self.Implementation["UseDiscreteDopplerSpectrum"]:
            if gauss is None:
                ctx.new_realization_seeded(nl, self._seed, self._count * nl)
            else:
                ctx.new_realization_gauss(np.asarray(gauss).reshape(nl, 2 * ctx.n_doppler_shifts + 1, ctx.T))
        elif doppler_u is None:
            ctx.new_realization_seeded(nl, self._seed, self._count * nl)
        else:
            du = np.asarray(doppler_u).reshape(nl, ctx.T, self.Nr["Paths"])
            pu = np.asarray(phase_u).reshape(nl, ctx.T, self.Nr["Paths"])
            ctx.new_realization(np.stack([x.reshape(-1, order="F") for x in du]), np.stack([x.reshape(-1, order="F") for x in pu]))
        self._count += 1
        hs = [ctx.impulse_response(b) for b in range(nl)]
        self.ImpulseResponse = hs[0] if siso else self._as_cells(hs)

    def Convolution(self, signal):
        """FF.m:253-274: r(:,iRx) = sum_iTx H{iRx,iTx} * s(:,iTx) with the banded H applied on the GPU."""
        s = np.asarray(signal)
        nT, nR = self.Nr["txAntennas"], self.Nr["rxAntennas"]
        if nT == 1 and nR == 1:
            return self._device().convolve(s, 0)
        s = s.reshape(s.shape[0], nT)
        out = np.zeros((s.shape[0], nR), dtype=complex)
        for tx in range(nT):
            for rx in range(nR):
                out[:, rx] += self._device().convolve(s[:, tx], rx + nR * tx)
        return out

    def GetConvolutionMatrix(self):
        """FF.m:276-295: nRx x nTx cell of sparse N x N convolution matrices."""
        nT, nR = self.Nr["txAntennas"], self.Nr["rxAntennas"]
        return [[self._device().convolution_matrix(rx + nR * tx) for tx in range(nT)] for rx in range(nR)]

    # ------------------------------------------------------------------ statistics (host, setup-time)
    def GetTimeCorrelation(self):
        """FF.m:321-340."""
        N, dt, fD = self.Nr["SamplesTotal"], self.PHY["dt"], self.PHY["MaximumDopplerShift"]
        time = (np.arange(2 * N - 1) - (N - 1)) * dt
        if self.PHY["DopplerModel"] in ("Jakes", "Discrete-Jakes"):                      # FF.m:332-336
            return j0(np.pi * 2 * fD * time), time
        return np.sinc(2 * fD * time), time

    def GetMeanDelay(self):
        p = self.Implementation["PowerDelayProfileNormalized"]
        return float(np.sum(np.arange(len(p)) * self.PHY["dt"] * p))

    def GetRmsDelaySpread(self):
        p = self.Implementation["PowerDelayProfileNormalized"]
        tau = np.arange(len(p)) * self.PHY["dt"]
        return float(np.sqrt(np.sum(tau ** 2 * p) - self.GetMeanDelay() ** 2))
