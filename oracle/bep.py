"""Oracle restatement of Theory/BitErrorProbabilityDoublyFlatRayleigh.m: closed-form bit error
probability of a Gray-mapped constellation in a doubly-flat Rayleigh channel with perfect CSI.
Used as a statistical known answer.  Test infrastructure only -- see oracle/__init__.py."""
import numpy as np


def _ratio_cdf(Ey2, Eh2, Eyh, zR, zI):
    """GaussianRatioCDF, BEP.m:80-130."""
    a, b = Eyh / Eh2, Ey2 / Eh2
    zR, zI = np.asarray(zR, dtype=float), np.asarray(zI, dtype=float)
    out = np.full(zR.shape, np.nan)
    i0 = (zR == -np.inf) | (zI == -np.inf)
    i1 = (zR == np.inf) & (zI == np.inf)
    ire = (zI == np.inf) & np.isfinite(zR)
    iim = (zR == np.inf) & np.isfinite(zI)
    inn = np.isfinite(zR) & np.isfinite(zI)
    c = b - abs(a) ** 2
    out[i0], out[i1] = 0.0, 1.0
    out[ire] = 0.5 - (a.real - zR[ire]) / (2 * np.sqrt((a.real - zR[ire]) ** 2 + c))          # :98-100
    out[iim] = 0.5 - (a.imag - zI[iim]) / (2 * np.sqrt((a.imag - zI[iim]) ** 2 + c))          # :102-104
    r, i = zR[inn] - a.real, zI[inn] - a.imag
    out[inn] = (0.25 + r * (2 * np.arctan(i / np.sqrt(r ** 2 + c)) + np.pi) / (4 * np.pi * np.sqrt(r ** 2 + c))
                + i * (2 * np.arctan(r / np.sqrt(i ** 2 + c)) + np.pi) / (4 * np.pi * np.sqrt(i ** 2 + c)))   # :109-119
    return out


def bit_error_probability_doubly_flat_rayleigh(SNR_dB, SymbolMapping, BitMapping):
    """BEP.m:10-51."""
    s = np.asarray(SymbolMapping, dtype=complex).reshape(-1)
    bm = np.asarray(BitMapping)
    half = np.min(np.abs(s.real))                                                             # :20
    lo_r, hi_r, lo_i, hi_i = s.real - half, s.real + half, s.imag - half, s.imag + half       # :21-25
    lo_r[s.real == s.real.min()] = -np.inf; hi_r[s.real == s.real.max()] = np.inf             # :26-29
    lo_i[s.imag == s.imag.min()] = -np.inf; hi_i[s.imag == s.imag.max()] = np.inf
    out = []
    for snr in np.atleast_1d(SNR_dB):
        Pn = 10 ** (-snr / 10)
        Pm = np.zeros((len(s), len(s)))
        for k, x in enumerate(s):                                                             # :35-41
            args = (abs(x) ** 2 + Pn, 1.0, x)
            Pm[:, k] = (_ratio_cdf(*args, hi_r, hi_i) + _ratio_cdf(*args, lo_r, lo_i)
                        - _ratio_cdf(*args, lo_r, hi_i) - _ratio_cdf(*args, hi_r, lo_i))      # :67-72
        e = np.zeros((2, bm.shape[1]))
        for ib in range(bm.shape[1]):                                                         # :43-48
            for v in (0, 1):
                idx = bm[:, ib] == v
                e[v, ib] = np.mean(np.sum(Pm[np.ix_(~idx, idx)], axis=0))
        out.append(e.mean())
    return np.array(out)
