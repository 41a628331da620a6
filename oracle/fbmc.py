"""Oracle restatement of +Modulation/FBMC.m for the path the reference scripts exercise:
'Hermite-OQAM', polyphase implementation, complex transmit signal.
Test infrastructure only -- see oracle/__init__.py."""
import numpy as np


def _hermite_h(n, x):
    """FBMC.m:685-706 (Hermite polynomials, literal coefficients)."""
    if n == 0:
        return np.ones_like(x)
    if n == 4:
        return 12 + (-48) * x**2 + 16 * x**4
    if n == 8:
        return 1680 + (-13440) * x**2 + 13440 * x**4 + (-3584) * x**6 + 256 * x**8
    if n == 12:
        return (665280 + (-7983360) * x**2 + 13305600 * x**4 + (-7096320) * x**6
                + 1520640 * x**8 + (-135168) * x**10 + 4096 * x**12)
    if n == 16:
        return (518918400 + (-8302694400) * x**2 + 19372953600 * x**4
                + (-15498362880) * x**6 + 5535129600 * x**8 + (-984023040) * x**10
                + 89456640 * x**12 + (-3932160) * x**14 + 65536 * x**16)
    if n == 20:
        return (670442572800 + (-13408851456000) * x**2 + 40226554368000 * x**4
                + (-42908324659200) * x**6 + 21454162329600 * x**8
                + (-5721109954560) * x**10 + 866834841600 * x**12
                + (-76205260800) * x**14 + 3810263040 * x**16
                + (-99614720) * x**18 + 1048576 * x**20)
    raise ValueError(n)


def prototype_filter_hermite(T0, dt, OF):
    """FBMC.m:629-647.  t_filter = -(OF*T0):dt:(OF*T0-dt)."""
    n = int(round(2 * OF * T0 / dt))
    t = -(OF * T0) + np.arange(n) * dt
    arg = t / (T0 / np.sqrt(2))
    gauss = np.exp(-np.pi * arg**2)
    coeff = {0: 1.412692577, 4: -3.0145e-3, 8: -8.8041e-6, 12: -2.2611e-9,
             16: -4.4570e-15, 20: 1.8633e-16}
    p = np.zeros(n)
    for order, Hk in coeff.items():
        p = p + (1 / np.sqrt(T0) * _hermite_h(order, np.sqrt(2 * np.pi) * arg) * gauss) * Hk
    return p / np.sqrt(np.sum(np.abs(p) ** 2) * dt)


class FBMC:
    """FBMC.m:26-160 (constructor + SetDependentParameters)."""

    def __init__(self, Subcarriers, MCSymbols, SubcarrierSpacing, SamplingRate,
                 IntermediateFrequency, TransmitRealSignal, Method, OverlappingFactor,
                 InitialPhaseShift, UsePolyphase):
        if Method != "Hermite-OQAM" or not UsePolyphase or TransmitRealSignal:
            raise NotImplementedError("oracle covers Hermite-OQAM / polyphase / complex only")
        self.Nr = {"Subcarriers": int(Subcarriers), "MCSymbols": int(MCSymbols)}
        self.PHY = {"SubcarrierSpacing": float(SubcarrierSpacing),
                    "SamplingRate": float(SamplingRate),
                    "IntermediateFrequency": float(IntermediateFrequency),
                    "TransmitRealSignal": False}
        self.Method = Method
        self.PrototypeFilter = {"OverlappingFactor": int(OverlappingFactor)}
        self.Implementation = {"InitialPhaseShift": InitialPhaseShift, "UsePolyphase": True}
        self._set_dependent()

    def _set_dependent(self):
        PHY, Nr, Imp, PF = self.PHY, self.Nr, self.Implementation, self.PrototypeFilter
        fs, F = PHY["SamplingRate"], PHY["SubcarrierSpacing"]
        if (fs / (2 * F)) % 1 != 0:                                           # :65-69
            F = fs / (2 * round(fs / (2 * F)))
            PHY["SubcarrierSpacing"] = F
        if (PHY["IntermediateFrequency"] / F) % 1 != 0:                       # :71-75
            PHY["IntermediateFrequency"] = round(PHY["IntermediateFrequency"] / F) * F
        if fs < Nr["Subcarriers"] * F:                                        # :77-79
            raise ValueError("Sampling Rate must be higher")
        PHY["dt"] = 1 / fs                                                    # :82
        Imp["TimeSpacing"] = int(round(fs / (2 * F)))                         # :87
        PHY["TimeSpacing"] = Imp["TimeSpacing"] * PHY["dt"]                   # :88
        Imp["FrequencySpacing"] = PF["OverlappingFactor"]                     # :89
        PF["TimeDomain"] = prototype_filter_hermite(PHY["TimeSpacing"] * 2, PHY["dt"],
                                                    PF["OverlappingFactor"] / 2)   # :90
        Np = len(PF["TimeDomain"])
        Nr["SamplesPrototypeFilter"] = Np                                     # :127
        Nr["SamplesTotal"] = Np + (Nr["MCSymbols"] - 1) * Imp["TimeSpacing"]  # :128
        L, K = Nr["Subcarriers"], Nr["MCSymbols"]
        k, l = np.meshgrid(np.arange(K), np.arange(L))                        # :138
        Imp["PhaseShift"] = np.exp(1j * np.pi / 2 * (l + k)) * np.exp(1j * Imp["InitialPhaseShift"])  # :139
        # :142-149 -- 0-based sample index of every (filter tap, symbol)
        Imp["IndexNumberAfterIFFT"] = (np.arange(Np)[:, None]
                                       + np.arange(K)[None, :] * Imp["TimeSpacing"])
        Imp["FFTSize"] = int(round(Np / Imp["FrequencySpacing"]))             # :152
        Imp["IntermediateFrequency"] = int(round(PHY["IntermediateFrequency"] / F))   # :153
        rows = np.zeros(Imp["FFTSize"], dtype=bool)
        rows[:L] = True
        Imp["IndexPolyphaseRows"] = np.roll(rows, Imp["IntermediateFrequency"])        # :154-156
        Imp["NormalizationFactor"] = np.sqrt(fs**2 / F**2 * PHY["TimeSpacing"] / L)   # :159

    # ------------------------------------------------------------------ modem
    def Modulation(self, DataSymbols):
        """FBMC.m:255-268 (polyphase branch).  DataSymbols: L x K.  Returns N samples."""
        Imp, Nr = self.Implementation, self.Nr
        K = Nr["MCSymbols"]
        X = np.zeros((Imp["FFTSize"], K), dtype=complex)
        X[Imp["IndexPolyphaseRows"], :] = (np.asarray(DataSymbols) * Imp["PhaseShift"]
                                           * Imp["NormalizationFactor"])      # :263
        blocks = np.tile(np.fft.ifft(X, axis=0), (Imp["FrequencySpacing"], 1)) \
            * self.PrototypeFilter["TimeDomain"][:, None]                    # :267
        s = np.zeros(Nr["SamplesTotal"], dtype=complex)
        np.add.at(s, Imp["IndexNumberAfterIFFT"].reshape(-1), blocks.reshape(-1))   # :267-268
        return s

    def Demodulation(self, ReceivedSignal):
        """FBMC.m:287-302 (polyphase branch).  Returns L x K."""
        Imp, Nr = self.Implementation, self.Nr
        K = Nr["MCSymbols"]
        r = np.asarray(ReceivedSignal).reshape(-1)
        seg = r[Imp["IndexNumberAfterIFFT"]] * self.PrototypeFilter["TimeDomain"][:, None]   # :294,297
        folded = seg.reshape(Imp["FrequencySpacing"], Imp["FFTSize"], K).sum(axis=0)        # :298
        Y = np.fft.fft(folded, axis=0)
        return (Y[Imp["IndexPolyphaseRows"], :] * np.conj(Imp["PhaseShift"])
                / (Imp["NormalizationFactor"] * self.PHY["SubcarrierSpacing"]))            # :302

    # ------------------------------------------------------- matrix description
    def GetTXMatrix(self):
        """FBMC.m:318-342: G with s = G*x(:)."""
        Nr, Imp = self.Nr, self.Implementation
        L, K, N = Nr["Subcarriers"], Nr["MCSymbols"], Nr["SamplesTotal"]
        tmp = np.zeros((N, L), dtype=complex)
        x = np.zeros((L, K))
        for l in range(L):                                                   # :330-334
            x[l, 0] = 1
            tmp[:, l] = self.Modulation(x)
            x[l, 0] = 0
        G = np.zeros((N, L * K), dtype=complex)
        for k in range(K):                                                   # :335-337
            G[:, k * L:(k + 1) * L] = np.roll(tmp, k * Imp["TimeSpacing"], axis=0) * (1j ** k)
        return G

    def GetRXMatrix(self):
        """FBMC.m:343-354: Q with y = Q*r."""
        return self.GetTXMatrix().conj().T * (self.Nr["Subcarriers"]
                                              / (self.PHY["SamplingRate"] * self.PHY["TimeSpacing"]))

    def GetInterferenceMatrix(self):
        """FBMC.m:390-400."""
        L, K = self.Nr["Subcarriers"], self.Nr["MCSymbols"]
        x = np.zeros((L, K))
        x[0, 0] = 1
        Y11 = self.Demodulation(self.Modulation(x))
        k_all, l_all = np.meshgrid(np.arange(K), np.arange(L))
        Y11 = Y11 * (np.exp(1j * np.pi / 2 * (l_all + k_all)) * np.exp(-1j * np.pi * k_all * (l_all / 2)))
        top = np.hstack([Y11[:0:-1, :0:-1], Y11[:0:-1, :]])
        bot = np.hstack([Y11[:, :0:-1], Y11])
        return np.vstack([top, bot])

    def GetFBMCMatrix(self, FastCalculation=True):
        """FBMC.m:355-388: D0 with y = D0*x for a flat channel."""
        L, K = self.Nr["Subcarriers"], self.Nr["MCSymbols"]
        LK = L * K
        if FastCalculation:
            IM = self.GetInterferenceMatrix()
            sym_i, sub_i = np.meshgrid(np.arange(1, K + 1), np.arange(1, L + 1))
            sub = sub_i.flatten(order="F")
            sym = sym_i.flatten(order="F")
            d_sub = sub[:, None] - sub[None, :]                              # :367
            d_sym = sym[:, None] - sym[None, :]                              # :368
            idx_sub = np.repeat(sub[:, None], LK, axis=1) - 1                # :369
            D0 = IM[d_sub + L - 1, d_sym + K - 1]                            # :370
            D0 = (D0 * np.exp(-1j * np.pi / 2 * (d_sub + d_sym))
                  * np.exp(-1j * 2 * np.pi * (self.PHY["TimeSpacing"] * self.PHY["SubcarrierSpacing"])
                           * d_sym * (idx_sub + d_sub / 2)))                 # :376
            return D0
        D0 = np.zeros((LK, LK), dtype=complex)                               # :380-386
        imp = np.zeros(LK)
        for i in range(LK):
            imp[i] = 1
            D0[:, i] = self.Demodulation(self.Modulation(imp.reshape(L, K, order="F"))).reshape(-1, order="F")
            imp[i] = 0
        return D0
