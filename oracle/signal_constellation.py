"""Oracle restatement of +Modulation/SignalConstellation.m (ctor, Bit2Symbol, Symbol2Bit,
SymbolQuantization).  Test infrastructure only -- see oracle/__init__.py."""
import numpy as np


def bi2de(bits):
    """Communications-Toolbox bi2de default ('right-msb'): first column is the LSB
    (SignalConstellation.m:64,79)."""
    bits = np.asarray(bits).astype(np.int64)
    w = (1 << np.arange(bits.shape[1], dtype=np.int64))
    return bits @ w


def _gray_atom(n_levels):
    """Gray-coded per-axis bit table, SignalConstellation.m:36-40 / 51-55."""
    nb = int(round(np.log2(n_levels)))
    atom = np.zeros((n_levels, nb), dtype=np.int64)
    atom[: n_levels // 2, 0] = 1
    for i in range(1, nb):
        tmp = atom[0::2, i - 1]
        atom[:, i] = np.concatenate([tmp, tmp[::-1]])
    return atom


class SignalConstellation:
    """SignalConstellation.m:24-74."""

    def __init__(self, ModulationOrder, Method):
        self.ModulationOrder = int(ModulationOrder)
        self.Method = Method
        M = self.ModulationOrder
        nb = int(round(np.log2(M)))
        if Method == "QAM":
            ms = int(round(np.sqrt(M)))
            atom = _gray_atom(ms)                               # :36-40
            IQ = 2 * np.arange(1, ms + 1) - ms - 1              # :41
            I_rep, Q_rep = np.meshgrid(IQ, IQ)                  # :42
            I_col = I_rep.flatten(order="F")
            Q_col = Q_rep.flatten(order="F")
            sym = I_col + 1j * Q_col                            # :43
            sym = sym / np.sqrt(np.mean(np.abs(sym) ** 2))      # :44
            bitmap = np.zeros((M, nb), dtype=np.int64)          # :45
            for x in IQ:                                        # :46-49
                bitmap[np.ix_(np.flatnonzero(I_col == x), np.arange(1, nb, 2))] = atom
                bitmap[np.ix_(np.flatnonzero(Q_col == x), np.arange(0, nb, 2))] = atom
        elif Method == "PAM":
            bitmap = _gray_atom(M)                              # :51-55
            sym = (2 * np.arange(1, M + 1) - M - 1).astype(np.float64)   # :56
            sym = sym / np.sqrt(np.mean(np.abs(sym) ** 2))      # :57
        else:
            raise ValueError("Signal constellation method must be QAM or PAM!")
        order = np.argsort(bi2de(bitmap), kind="stable")        # :64
        self.SymbolMapping = sym[order]                         # :65
        self.BitMapping = bitmap[order, :]                      # :66

    def Bit2Symbol(self, BinaryStream):
        """SignalConstellation.m:76-81 (column vector in, column vector out)."""
        nb = self.BitMapping.shape[1]
        b = np.asarray(BinaryStream).reshape(-1, nb)            # reshape(.,nb,[])'
        return self.SymbolMapping[bi2de(b)]

    def _nearest(self, x):
        x = np.asarray(x).reshape(-1)
        d = np.abs(x[:, None] - self.SymbolMapping[None, :])    # :88 / :98
        return np.argmin(d, axis=1)                             # first index on ties

    def Symbol2Bit(self, EstimatedDataSymbols):
        """SignalConstellation.m:83-91; output is symbol-major, LSB first."""
        return self.BitMapping[self._nearest(EstimatedDataSymbols), :].reshape(-1)

    def SymbolQuantization(self, EstimatedDataSymbols):
        """SignalConstellation.m:93-101."""
        return self.SymbolMapping[self._nearest(EstimatedDataSymbols)]
