"""Oracle restatement of +Modulation/OFDM.m (complex transmit signal path).
Test infrastructure only -- see oracle/__init__.py."""
import numpy as np


class OFDM:
    """OFDM.m:23-88."""

    def __init__(self, Subcarriers, MCSymbols, SubcarrierSpacing, SamplingRate,
                 IntermediateFrequency, TransmitRealSignal, CyclicPrefixLength,
                 ZeroGuardTimeLength):
        if TransmitRealSignal:
            raise NotImplementedError("oracle covers the complex transmit signal only")
        self.Nr = {"Subcarriers": int(Subcarriers), "MCSymbols": int(MCSymbols)}
        self.PHY = {"SubcarrierSpacing": float(SubcarrierSpacing), "SamplingRate": float(SamplingRate),
                    "IntermediateFrequency": float(IntermediateFrequency), "TransmitRealSignal": False,
                    "CyclicPrefixLength": float(CyclicPrefixLength),
                    "ZeroGuardTimeLength": float(ZeroGuardTimeLength)}
        self.Implementation = {}
        self._set_dependent()

    def _set_dependent(self):
        PHY, Nr, Imp = self.PHY, self.Nr, self.Implementation
        fs = PHY["SamplingRate"]
        if (round(fs / PHY["SubcarrierSpacing"] * 1e5) / 1e5) % 1 != 0:                     # :57-61
            PHY["SubcarrierSpacing"] = fs / round(fs / PHY["SubcarrierSpacing"])
        F = PHY["SubcarrierSpacing"]
        if (round(PHY["IntermediateFrequency"] / F * 1e5) / 1e5) % 1 != 0:                  # :63-67
            PHY["IntermediateFrequency"] = round(PHY["IntermediateFrequency"] / F) * F
        if fs < Nr["Subcarriers"] * F:                                                      # :69-71
            raise ValueError("Sampling theorem is not fullfilled")
        if abs((round(PHY["CyclicPrefixLength"] * fs * 1e5) / 1e5) % 1) != 0:               # :73-77
            PHY["CyclicPrefixLength"] = _matlab_round(PHY["CyclicPrefixLength"] * fs) / fs
        Imp["CyclicPrefix"] = _matlab_round(PHY["CyclicPrefixLength"] * fs)                 # :79
        Imp["ZeroGuardSamples"] = _matlab_round(PHY["ZeroGuardTimeLength"] * fs)            # :80
        Imp["TimeSpacing"] = _matlab_round(fs / F) + Imp["CyclicPrefix"]                    # :81
        Imp["FFTSize"] = _matlab_round(fs / F)                                              # :82
        Imp["IntermediateFrequency"] = _matlab_round(PHY["IntermediateFrequency"] / F)      # :83
        Imp["NormalizationFactor"] = np.sqrt(fs**2 / F**2 / Nr["Subcarriers"])              # :84
        PHY["dt"] = 1 / fs                                                                  # :85
        PHY["TimeSpacing"] = Imp["TimeSpacing"] * PHY["dt"]                                 # :86
        Nr["SamplesTotal"] = Nr["MCSymbols"] * Imp["TimeSpacing"] + 2 * Imp["ZeroGuardSamples"]   # :87

    def Modulation(self, DataSymbols):
        """OFDM.m:153-165.  DataSymbols: L x K; returns N samples."""
        Imp, Nr = self.Implementation, self.Nr
        L, K = Nr["Subcarriers"], Nr["MCSymbols"]
        X = np.zeros((Imp["FFTSize"], K), dtype=complex)
        X[Imp["IntermediateFrequency"]:Imp["IntermediateFrequency"] + L, :] = \
            np.asarray(DataSymbols) * Imp["NormalizationFactor"]                            # :159
        noCP = np.fft.ifft(X, axis=0)                                                       # :163
        cp = Imp["CyclicPrefix"]
        withCP = np.vstack([noCP[Imp["FFTSize"] - cp:, :], noCP])                           # :164
        zg = np.zeros(Imp["ZeroGuardSamples"], dtype=complex)
        return np.concatenate([zg, withCP.reshape(-1, order="F"), zg])

    def Demodulation(self, ReceivedSignal):
        """OFDM.m:167-181.  Returns L x K."""
        Imp, Nr = self.Implementation, self.Nr
        L, K = Nr["Subcarriers"], Nr["MCSymbols"]
        r = np.asarray(ReceivedSignal).reshape(-1)
        zg = Imp["ZeroGuardSamples"]
        core = r[zg:len(r) - zg].reshape(Imp["TimeSpacing"], K, order="F")                  # :173
        Y = np.fft.fft(core[Imp["CyclicPrefix"]:, :], axis=0)                               # :174
        return Y[Imp["IntermediateFrequency"]:Imp["IntermediateFrequency"] + L, :] / Imp["NormalizationFactor"]

    def GetTXMatrix(self):
        """OFDM.m:184-203."""
        Nr, Imp = self.Nr, self.Implementation
        L, K, N = Nr["Subcarriers"], Nr["MCSymbols"], Nr["SamplesTotal"]
        tmp = np.zeros((N, L), dtype=complex)
        x = np.zeros((L, K))
        for l in range(L):                                                                  # :195-199
            x[l, 0] = 1
            tmp[:, l] = self.Modulation(x)
            x[l, 0] = 0
        G = np.zeros((N, L * K), dtype=complex)
        for k in range(K):                                                                  # :200-202
            G[:, k * L:(k + 1) * L] = np.roll(tmp, k * Imp["TimeSpacing"], axis=0)
        return G

    def GetRXMatrix(self):
        """OFDM.m:205-218: scaled G^H with the cyclic-prefix sample columns zeroed."""
        Nr, Imp = self.Nr, self.Implementation
        R = self.GetTXMatrix().conj().T * (Nr["Subcarriers"] * self.PHY["SubcarrierSpacing"]
                                           / self.PHY["SamplingRate"])                      # :214
        idx = (Imp["ZeroGuardSamples"] + np.arange(Imp["CyclicPrefix"])[:, None]
               + np.arange(Nr["MCSymbols"])[None, :] * Imp["TimeSpacing"])                  # :216
        R[:, idx.reshape(-1)] = 0                                                           # :217
        return R


def _matlab_round(x):
    """MATLAB round(): half away from zero."""
    return int(np.floor(abs(x) + 0.5) * (1 if x >= 0 else -1))
