"""Host mirror of the reference's +Modulation package (same class / method / property names).

`Modulation.FBMC`, `Modulation.OFDM`: the constructors reproduce SetDependentParameters
(FBMC.m:61-160, OFDM.m:53-88); the transmit / receive matrices are written in closed form (the
reference obtains them by calling its IFFT modulator once per subcarrier, FBMC.m:330-337);
`Modulation` / `Demodulation` run on the GPU in their FFT form (polyphase IFFT + prototype filter + overlap-add,
FBMC.m:255-302; CP-OFDM, OFDM.m:153-181: hand-written mixed-radix FFT kernels, chest_set_modem /
chest_modulate_fft / chest_demodulate_fft); `ModulationMatrix` / `DemodulationMatrix` keep the matrix form
s = G x(:), y = Q' r (FBMC.m:319-320,344-345; OFDM.m:185-186,206-207), the form the DS.m loop body uses.
`Modulation.SignalConstellation`: tables of SC.m:24-74; the nearest-neighbour demapping used on
the hot path lives in the CUDA library (kernels.cuh, demap_word)."""
import sys

import numpy as np

from .context import DeviceContext


def _disp(msg):
    """The reference disp()s a note when it adjusts an inconsistent parameter; keep stdout clean."""
    print(msg, file=sys.stderr)


def _round_half_away(x):
    return int(np.floor(abs(x) + 0.5) * (1 if x >= 0 else -1))


# ----------------------------------------------------------------------------- constellation
class SignalConstellation:
    """Modulation.SignalConstellation(ModulationOrder, 'QAM'|'PAM')  (SC.m:24-74)."""

    def __init__(self, ModulationOrder, Method):
        M = int(ModulationOrder)
        self.ModulationOrder, self.Method = M, Method
        if Method == "PAM":
            gray = self._axis_bits(M)
            sym = (2.0 * np.arange(1, M + 1) - M - 1).astype(np.complex128)
        elif Method == "QAM":
            ms = int(round(np.sqrt(M)))
            if ms * ms != M:
                raise ValueError("QAM order must be a square")
            ax = self._axis_bits(ms)
            nb = ax.shape[1]
            lev = 2.0 * np.arange(1, ms + 1) - ms - 1
            # column-major grid: index = q + ms*i ; odd bit columns follow the I level's rows,
            # even bit columns the Q level's rows (SC.m:46-49)
            ii, qq = np.divmod(np.arange(M), ms)
            sym = lev[ii] + 1j * lev[qq]
            gray = np.zeros((M, 2 * nb), dtype=np.int64)
            gray[:, 1::2] = ax[qq, :]
            gray[:, 0::2] = ax[ii, :]
        else:
            raise ValueError("Signal constellation method must be QAM or PAM!")
        sym = sym / np.sqrt(np.mean(np.abs(sym) ** 2))
        words = gray @ (1 << np.arange(gray.shape[1]))          # bi2de, first column = LSB (SC.m:64)
        order = np.argsort(words, kind="stable")
        self.SymbolMapping = sym[order]
        self.BitMapping = gray[order, :].astype(np.uint8)
        if Method == "PAM":
            self.SymbolMapping = self.SymbolMapping.real.astype(np.float64)

    @staticmethod
    def _axis_bits(n):
        """Gray labels of one amplitude axis (SC.m:36-40,51-55)."""
        nb = int(round(np.log2(n)))
        t = np.zeros((n, nb), dtype=np.int64)
        t[: n // 2, 0] = 1
        for c in range(1, nb):
            half = t[0::2, c - 1]
            t[:, c] = np.concatenate([half, half[::-1]])
        return t

    def Bit2Symbol(self, BinaryStream):
        """SC.m:76-81."""
        nb = self.BitMapping.shape[1]
        b = np.asarray(BinaryStream).astype(np.int64).reshape(-1, nb)
        return self.SymbolMapping[b @ (1 << np.arange(nb))]

    def _nearest(self, x):
        x = np.asarray(x).reshape(-1)
        return np.argmin(np.abs(x[:, None] - self.SymbolMapping[None, :]), axis=1)

    def Symbol2Bit(self, EstimatedDataSymbols):
        """SC.m:83-91 (host convenience; the hot path demaps on the device)."""
        return self.BitMapping[self._nearest(EstimatedDataSymbols), :].reshape(-1)

    def SymbolQuantization(self, EstimatedDataSymbols):
        """SC.m:93-101."""
        return self.SymbolMapping[self._nearest(EstimatedDataSymbols)]


# ----------------------------------------------------------------------------- shared modem plumbing
class _Modem:
    _wf = "F"

    def _device(self):
        """Context with the dense G / Q uploaded (matrix-form modem)."""
        if getattr(self, "_ctx", None) is None:
            self._ctx = DeviceContext()
            self._ctx.set_waveform(self._wf, self.GetTXMatrix(), self.GetRXMatrix().conj().T)
        return self._ctx

    def _device_fft(self):
        """Context with only the FFT modem's tables (no N x LK matrices)."""
        if getattr(self, "_ctx_fft", None) is None:
            self._ctx_fft = DeviceContext()
            self._set_modem(self._ctx_fft)
        return self._ctx_fft

    def _check_x(self, DataSymbols):
        x = np.asarray(DataSymbols)
        L, K = self.Nr["Subcarriers"], self.Nr["MCSymbols"]
        if x.shape[:2] != (L, K):
            raise ValueError("DataSymbols must be Subcarriers x MCSymbols")
        return x, x.reshape(L * K, -1, order="F")

    def Modulation(self, DataSymbols):
        """FBMC.m:255-268 / OFDM.m:153-165 on the GPU (FFT form).  DataSymbols: L x K (or several stacked along a
        3rd axis, as SimpleVersion_DoublyFlat.m:118 does)."""
        x, flat = self._check_x(DataSymbols)
        s = self._device_fft().modulate_fft(self._wf, flat)
        return s[:, 0] if x.ndim == 2 else s

    def Demodulation(self, ReceivedSignal):
        """FBMC.m:287-302 / OFDM.m:167-181 on the GPU (FFT form)."""
        r = np.asarray(ReceivedSignal)
        L, K = self.Nr["Subcarriers"], self.Nr["MCSymbols"]
        if r.shape[0] != self.Nr["SamplesTotal"]:
            raise ValueError("ReceivedSignal must have Nr.SamplesTotal rows")
        y = self._device_fft().demodulate_fft(self._wf, r.reshape(r.shape[0], -1))
        return y[:, 0].reshape(L, K, order="F") if r.ndim == 1 else y.reshape(L, K, -1, order="F")

    def ModulationMatrix(self, DataSymbols):
        """s = G*x(:) on the GPU (the identity the reference states at FBMC.m:319-320 / OFDM.m:185-186)."""
        x, flat = self._check_x(DataSymbols)
        s = self._device().modulate(self._wf, flat)
        return s[:, 0] if x.ndim == 2 else s

    def DemodulationMatrix(self, ReceivedSignal):
        """reshape(Q'*r, L, K) on the GPU (FBMC.m:344-345 / OFDM.m:206-207)."""
        r = np.asarray(ReceivedSignal)
        L, K = self.Nr["Subcarriers"], self.Nr["MCSymbols"]
        y = self._device().demodulate(self._wf, r.reshape(r.shape[0], -1))
        return y[:, 0].reshape(L, K, order="F") if r.ndim == 1 else y.reshape(L, K, -1, order="F")


# ----------------------------------------------------------------------------- FBMC
def _hermite_prototype(T0, dt, OF):
    """PrototypeFilter_Hermite (FBMC.m:629-647), Horner form of the even Hermite polynomials."""
    n = int(round(2 * OF * T0 / dt))
    t = -(OF * T0) + np.arange(n) * dt
    u = t / (T0 / np.sqrt(2))
    z = (np.sqrt(2 * np.pi) * u) ** 2
    herm = {
        0: [1.0],
        4: [12.0, -48.0, 16.0],
        8: [1680.0, -13440.0, 13440.0, -3584.0, 256.0],
        12: [665280.0, -7983360.0, 13305600.0, -7096320.0, 1520640.0, -135168.0, 4096.0],
        16: [518918400.0, -8302694400.0, 19372953600.0, -15498362880.0, 5535129600.0, -984023040.0,
             89456640.0, -3932160.0, 65536.0],
        20: [670442572800.0, -13408851456000.0, 40226554368000.0, -42908324659200.0, 21454162329600.0,
             -5721109954560.0, 866834841600.0, -76205260800.0, 3810263040.0, -99614720.0, 1048576.0],
    }
    weight = {0: 1.412692577, 4: -3.0145e-3, 8: -8.8041e-6, 12: -2.2611e-9, 16: -4.4570e-15, 20: 1.8633e-16}
    p = np.zeros(n)
    for order, coef in herm.items():
        poly = np.zeros(n)
        for ck in coef[::-1]:
            poly = poly * z + ck
        p += weight[order] * poly
    p *= np.exp(-np.pi * u ** 2) / np.sqrt(T0)
    return p / np.sqrt(np.sum(p ** 2) * dt)


class FBMC(_Modem):
    """Modulation.FBMC(Subcarriers, MCSymbols, SubcarrierSpacing, SamplingRate, IntermediateFrequency,
    TransmitRealSignal, Method, OverlappingFactor, InitialPhaseShift, UsePolyphase)  (FBMC.m:28-59)."""
    _wf = "F"

    def __init__(self, *args):
        if len(args) == 0:
            args = (12, 30, 15e3, 12 * 15e3, 0, False, "Hermite-OQAM", 8, 0, True)      # FBMC.m:43-52
        if len(args) != 10:
            raise ValueError("Number of input variables must be either 0 (default values) or 10")
        (L, K, F, fs, f_if, real_sig, method, OF, phi0, polyphase) = args
        if method != "Hermite-OQAM":
            raise NotImplementedError('Method (prototype filter) "%s" is not supported by this build' % method)
        if real_sig:
            raise NotImplementedError("TransmitRealSignal = true is not supported by this build")
        self.Method = method
        self.Nr = {"Subcarriers": int(L), "MCSymbols": int(K)}
        self.PHY = {"SubcarrierSpacing": float(F), "SamplingRate": float(fs), "IntermediateFrequency": float(f_if),
                    "TransmitRealSignal": False}
        self.PrototypeFilter = {"OverlappingFactor": int(OF)}
        self.Implementation = {"InitialPhaseShift": phi0, "UsePolyphase": bool(polyphase)}
        self.SetDependentParameters()

    def SetDependentParameters(self):
        PHY, Nr, Imp = self.PHY, self.Nr, self.Implementation
        fs = PHY["SamplingRate"]
        if (fs / (2 * PHY["SubcarrierSpacing"])) % 1 != 0:                                  # FBMC.m:65-69
            PHY["SubcarrierSpacing"] = fs / (2 * _round_half_away(fs / (2 * PHY["SubcarrierSpacing"])))
            _disp("Sampling Rate divided by (Subcarrier spacing times 2) must be must be an integer!")
        F = PHY["SubcarrierSpacing"]
        if (PHY["IntermediateFrequency"] / F) % 1 != 0:                                     # FBMC.m:71-75
            PHY["IntermediateFrequency"] = _round_half_away(PHY["IntermediateFrequency"] / F) * F
            _disp("The intermediate frequency must be a multiple of the subcarrier spacing!")
        if fs < Nr["Subcarriers"] * F:                                                      # FBMC.m:77-79
            raise ValueError("Sampling Rate must be higher: at least Number of Subcarriers times Subcarrier Spacing")
        PHY["dt"] = 1.0 / fs
        Imp["TimeSpacing"] = int(round(fs / (2 * F)))                                       # FBMC.m:87
        PHY["TimeSpacing"] = Imp["TimeSpacing"] * PHY["dt"]
        Imp["FrequencySpacing"] = self.PrototypeFilter["OverlappingFactor"]
        self.PrototypeFilter["TimeDomain"] = _hermite_prototype(PHY["TimeSpacing"] * 2, PHY["dt"],
                                                                self.PrototypeFilter["OverlappingFactor"] / 2)
        Np = len(self.PrototypeFilter["TimeDomain"])
        Nr["SamplesPrototypeFilter"] = Np
        Nr["SamplesTotal"] = Np + (Nr["MCSymbols"] - 1) * Imp["TimeSpacing"]                # FBMC.m:128
        Imp["FFTSize"] = int(round(Np / Imp["FrequencySpacing"]))                           # FBMC.m:152
        Imp["IntermediateFrequency"] = int(round(PHY["IntermediateFrequency"] / F))
        Imp["NormalizationFactor"] = np.sqrt(fs ** 2 / F ** 2 * PHY["TimeSpacing"] / Nr["Subcarriers"])
        self._ctx = None
        self._G = None

    def _set_modem(self, ctx):
        Nr, Imp = self.Nr, self.Implementation
        L, K, nfft = Nr["Subcarriers"], Nr["MCSymbols"], Imp["FFTSize"]
        bins = np.sort((Imp["IntermediateFrequency"] + np.arange(L)) % nfft)          # FBMC.m:154-156: ascending rows
        k, l = np.meshgrid(np.arange(K), np.arange(L))
        phase = np.exp(1j * np.pi / 2 * (l + k)) * np.exp(1j * Imp["InitialPhaseShift"])          # FBMC.m:139
        ctx.set_modem(self._wf, "fbmc", L, K, nfft, bins, Imp["TimeSpacing"], O=Imp["FrequencySpacing"],
                      prototype_filter=self.PrototypeFilter["TimeDomain"], phase_shift=phase,
                      normalization=Imp["NormalizationFactor"], subcarrier_spacing=self.PHY["SubcarrierSpacing"])

    def GetTXMatrix(self):
        """G (N x L*K): column (l,k) = shifted, modulated prototype filter.  Closed form of
        FBMC.m:318-342: an impulse on FFT bin b_l gives ifft = exp(j 2 pi b_l t / FFT) / FFT, tiled and
        windowed by the prototype filter; symbol k is that column delayed by k*TimeSpacing times j^k."""
        if self._G is not None:
            return self._G
        Nr, Imp = self.Nr, self.Implementation
        L, K, N, Np = Nr["Subcarriers"], Nr["MCSymbols"], Nr["SamplesTotal"], Nr["SamplesPrototypeFilter"]
        nfft, TS = Imp["FFTSize"], Imp["TimeSpacing"]
        # logical polyphase map rows are filled in ascending row order (FBMC.m:154-156,263)
        bins = np.sort((Imp["IntermediateFrequency"] + np.arange(L)) % nfft)
        t = np.arange(Np)
        base = (self.PrototypeFilter["TimeDomain"][:, None] * (Imp["NormalizationFactor"] / nfft)
                * np.exp(2j * np.pi * np.outer(t, bins) / nfft)
                * (np.exp(1j * np.pi / 2 * np.arange(L)) * np.exp(1j * Imp["InitialPhaseShift"]))[None, :])
        G = np.zeros((N, L * K), dtype=np.complex128)
        for k in range(K):
            G[k * TS:k * TS + Np, k * L:(k + 1) * L] = base * (1j ** k)
        self._G = G
        return G

    def GetRXMatrix(self):
        """FBMC.m:343-354."""
        return self.GetTXMatrix().conj().T * (self.Nr["Subcarriers"] / (self.PHY["SamplingRate"] * self.PHY["TimeSpacing"]))

    def GetFBMCMatrix(self, FastCalculation=True):
        """D0 with y = D0*x over a flat channel (FBMC.m:355-388).  The fast variant looks every entry
        up in the interference pattern of one impulse by (delta subcarrier, delta symbol), so entries
        with equal offsets are bit-identical -- the pilot precoders' tie handling relies on that."""
        if not FastCalculation:
            return self.GetRXMatrix() @ self.GetTXMatrix()
        L, K = self.Nr["Subcarriers"], self.Nr["MCSymbols"]
        IM = self.GetInterferenceMatrix()
        l = np.tile(np.arange(L), K)
        k = np.repeat(np.arange(K), L)
        dl, dk = l[:, None] - l[None, :], k[:, None] - k[None, :]
        D0 = IM[dl + L - 1, dk + K - 1]
        TF = self.PHY["TimeSpacing"] * self.PHY["SubcarrierSpacing"]
        return (D0 * np.exp(-1j * np.pi / 2 * (dl + dk))
                * np.exp(-1j * 2 * np.pi * TF * dk * (l[:, None] + dl / 2)))                  # FBMC.m:376

    def GetInterferenceMatrix(self):
        """FBMC.m:390-400."""
        L, K = self.Nr["Subcarriers"], self.Nr["MCSymbols"]
        Y = (self.GetRXMatrix() @ self.GetTXMatrix()[:, 0]).reshape(L, K, order="F")
        k_all, l_all = np.meshgrid(np.arange(K), np.arange(L))
        Y = Y * np.exp(1j * np.pi / 2 * (l_all + k_all)) * np.exp(-1j * np.pi * k_all * (l_all / 2))
        return np.vstack([np.hstack([Y[:0:-1, :0:-1], Y[:0:-1, :]]), np.hstack([Y[:, :0:-1], Y])])


# ----------------------------------------------------------------------------- OFDM
class OFDM(_Modem):
    """Modulation.OFDM(Subcarriers, MCSymbols, SubcarrierSpacing, SamplingRate, IntermediateFrequency,
    TransmitRealSignal, CyclicPrefixLength, ZeroGuardTimeLength)  (OFDM.m:23-51)."""
    _wf = "O"

    def __init__(self, *args):
        if len(args) == 0:
            args = (24, 14, 15e3, 15e3 * 24 * 14, 0, False, 1 / (14 * 15e3), 0)              # OFDM.m:36-43
        if len(args) != 8:
            raise ValueError("Number of input variables must be either 0 (default values) or 8")
        (L, K, F, fs, f_if, real_sig, cp, zg) = args
        if real_sig:
            raise NotImplementedError("GetTXMatrix is not supported for PHY.TransmitRealSignal == true!")
        self.Nr = {"Subcarriers": int(L), "MCSymbols": int(K)}
        self.PHY = {"SubcarrierSpacing": float(F), "SamplingRate": float(fs), "IntermediateFrequency": float(f_if),
                    "TransmitRealSignal": False, "CyclicPrefixLength": float(cp), "ZeroGuardTimeLength": float(zg)}
        self.Implementation = {}
        self.SetDependentParameters()

    def SetDependentParameters(self):
        PHY, Nr, Imp = self.PHY, self.Nr, self.Implementation
        fs = PHY["SamplingRate"]
        if (round(fs / PHY["SubcarrierSpacing"] * 1e5) / 1e5) % 1 != 0:                      # OFDM.m:57-61
            PHY["SubcarrierSpacing"] = fs / _round_half_away(fs / PHY["SubcarrierSpacing"])
            _disp("Sampling rate must be a multiple of the subcarrier spacing!")
        F = PHY["SubcarrierSpacing"]
        if (round(PHY["IntermediateFrequency"] / F * 1e5) / 1e5) % 1 != 0:                   # OFDM.m:63-67
            PHY["IntermediateFrequency"] = _round_half_away(PHY["IntermediateFrequency"] / F) * F
        if fs < Nr["Subcarriers"] * F:                                                       # OFDM.m:69-71
            raise ValueError("Sampling theorem is not fullfilled: sampling rate must be higher than the number "
                             "of subcarriers times subcarrier spacing")
        if abs((round(PHY["CyclicPrefixLength"] * fs * 1e5) / 1e5) % 1) != 0:                # OFDM.m:73-77
            PHY["CyclicPrefixLength"] = _round_half_away(PHY["CyclicPrefixLength"] * fs) / fs
            _disp("The length of the cyclic prefix times the sampling rate must be an integer!")
        Imp["CyclicPrefix"] = _round_half_away(PHY["CyclicPrefixLength"] * fs)
        Imp["ZeroGuardSamples"] = _round_half_away(PHY["ZeroGuardTimeLength"] * fs)
        Imp["FFTSize"] = _round_half_away(fs / F)
        Imp["TimeSpacing"] = Imp["FFTSize"] + Imp["CyclicPrefix"]
        Imp["IntermediateFrequency"] = _round_half_away(PHY["IntermediateFrequency"] / F)
        Imp["NormalizationFactor"] = np.sqrt(fs ** 2 / F ** 2 / Nr["Subcarriers"])
        PHY["dt"] = 1.0 / fs
        PHY["TimeSpacing"] = Imp["TimeSpacing"] * PHY["dt"]
        Nr["SamplesTotal"] = Nr["MCSymbols"] * Imp["TimeSpacing"] + 2 * Imp["ZeroGuardSamples"]
        self._ctx = None
        self._ctx_fft = None
        self._G = None

    def _set_modem(self, ctx):
        Nr, Imp = self.Nr, self.Implementation
        ctx.set_modem(self._wf, "ofdm", Nr["Subcarriers"], Nr["MCSymbols"], Imp["FFTSize"],
                      Imp["IntermediateFrequency"] + np.arange(Nr["Subcarriers"]), Imp["TimeSpacing"],
                      cp=Imp["CyclicPrefix"], zero_guard=Imp["ZeroGuardSamples"],
                      normalization=Imp["NormalizationFactor"], subcarrier_spacing=self.PHY["SubcarrierSpacing"])

    def GetTXMatrix(self):
        """G (N x L*K), closed form of OFDM.m:184-203: subcarrier l of symbol k is a complex
        exponential over FFTSize samples preceded by its cyclic prefix."""
        if self._G is not None:
            return self._G
        Nr, Imp = self.Nr, self.Implementation
        L, K, N = Nr["Subcarriers"], Nr["MCSymbols"], Nr["SamplesTotal"]
        nfft, cp, TS, zg = Imp["FFTSize"], Imp["CyclicPrefix"], Imp["TimeSpacing"], Imp["ZeroGuardSamples"]
        t = np.arange(TS) - cp
        base = (Imp["NormalizationFactor"] / nfft) * np.exp(
            2j * np.pi * np.outer(t % nfft, Imp["IntermediateFrequency"] + np.arange(L)) / nfft)
        G = np.zeros((N, L * K), dtype=np.complex128)
        for k in range(K):
            G[zg + k * TS:zg + (k + 1) * TS, k * L:(k + 1) * L] = base
        self._G = G
        return G

    def GetRXMatrix(self):
        """OFDM.m:205-218: scaled G' with the cyclic-prefix samples ignored."""
        Nr, Imp = self.Nr, self.Implementation
        R = self.GetTXMatrix().conj().T * (Nr["Subcarriers"] * self.PHY["SubcarrierSpacing"] / self.PHY["SamplingRate"])
        idx = (Imp["ZeroGuardSamples"] + np.arange(Imp["CyclicPrefix"])[:, None]
               + np.arange(Nr["MCSymbols"])[None, :] * Imp["TimeSpacing"])
        R[:, idx.reshape(-1)] = 0
        return R
