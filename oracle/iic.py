"""Oracle restatement of
+ChannelEstimation/ImaginaryInterferenceCancellationAtPilotPosition.m (both methods).
Test infrastructure only -- see oracle/__init__.py."""
import numpy as np
from scipy.linalg import hadamard


def _interference_matrix(FBMCMatrix, L, K):
    """IIC.m:47-51."""
    D0 = FBMCMatrix
    LK = L * K
    i11 = np.abs(D0[:, 0].reshape(L, K, order="F"))
    iE1 = np.abs(D0[:, L - 1].reshape(L, K, order="F"))
    i1E = np.abs(D0[:, LK - L].reshape(L, K, order="F"))
    iEE = np.abs(D0[:, LK - 1].reshape(L, K, order="F"))
    left = np.vstack([iEE, i1E[1:, :]])
    right = np.vstack([iE1[:, 1:], i11[1:, 1:]])
    return np.hstack([left, right])


# Tie rule of the interferer selection (IIC.m:72-73, 113-114).  The reference compares
#     abs(FBMCMatrix(pilot,:)) >= SortedInterferenceValues(N+1)
# with a plain >=.  When the FFT size equals the subcarrier count (DS.m default: fs = 24*15 kHz) the interference
# pattern holds groups of weights that are EQUAL in exact arithmetic (subcarrier offsets +l and L-l alias) and the
# (N+1)-th value falls inside such a group for the data-spreading scheme (N = 20: the 16th..31st largest values are all 0.036858; so does the 29th of the auxiliary scheme).
# In exact arithmetic >= takes the whole group; in floating point the members differ in their last bits (FFT and
# exp() rounding), so the literal comparison picks a rounding-noise-dependent subset -- a different one in MATLAB,
# in this NumPy restatement (12 of 16) and in the product's host mirror (14-16), i.e. the reference's behaviour
# is not defined beyond "some subset of the tie group".  TIE_RTOL > 0 restates the exact-arithmetic reading:
# values within TIE_RTOL (relative) below the threshold count as equal to it.  tie_rtol = 0 is the literal >=.
# At the paper geometry (DS.m:42-46) the threshold is the smallest member of its group and both readings agree.
TIE_RTOL = 1e-9


def _considered(FBMCMatrix, PilotMatrix, n_cancel, IM, tie_rtol=TIE_RTOL):
    """IIC.m:72-76 / 113-122: per-position pilot tags (-p for interferers of pilot p, +p at pilots)."""
    L, K = PilotMatrix.shape
    pm = PilotMatrix.reshape(-1, order="F")
    pil = np.flatnonzero(pm == 1)
    P = len(pil)
    srt = np.sort(np.abs(IM.reshape(-1)))[::-1]                                   # sort(...,'descend')
    tmp = np.abs(FBMCMatrix[pil, :]) >= srt[n_cancel] * (1.0 - tie_rtol)          # (NrCanceled+1)-th value
    ci = -(tmp * np.arange(1, P + 1)[:, None]).sum(axis=0).astype(np.int64)       # sum_p -(p) * mask_p
    ci[pil] = np.arange(1, P + 1)
    return tmp, ci


class ImaginaryInterferenceCancellationAtPilotPosition:
    """IIC.m:37-229."""

    def __init__(self, Method, PilotMatrix, FBMCMatrix, NrCanceledInterferersPerPilot,
                 PilotToDataPowerOffset, tie_rtol=TIE_RTOL):
        PilotMatrix = np.asarray(PilotMatrix)
        L, K = PilotMatrix.shape
        LK = L * K
        pm = PilotMatrix.reshape(-1, order="F")
        D0 = np.asarray(FBMCMatrix)
        IM = _interference_matrix(D0, L, K)
        pil = np.flatnonzero(pm == 1)
        P = len(pil)
        if Method == "Auxiliary":                                                 # :55-103
            dat = np.flatnonzero(pm == 0)
            aux = np.flatnonzero(pm == -1)
            nD, nA = len(dat), len(aux)
            pinv = np.linalg.pinv(D0[np.ix_(pil, aux)])                           # :60
            aux_p = pinv @ (np.eye(P) - D0[np.ix_(pil, pil)])                     # :61
            aux_d = -pinv @ D0[np.ix_(pil, dat)]                                  # :62
            C = np.zeros((LK, LK - nA), dtype=complex)                            # :65
            C[np.ix_(aux, np.arange(P))] = aux_p                                  # :66
            C[np.ix_(aux, np.arange(P, P + nD))] = aux_d                          # :67
            C[pil, np.arange(P)] = np.sqrt(PilotToDataPowerOffset)                # :68
            C[dat, P + np.arange(nD)] = 1.0                                       # :69
            if NrCanceledInterferersPerPilot > 0:                                 # :71-82
                _, ci = _considered(D0, PilotMatrix, NrCanceledInterferersPerPilot, IM, tie_rtol)
                idx = np.concatenate([ci[pil], ci[dat]])
                C[np.ix_(aux, np.flatnonzero(idx == 0))] = 0                      # :82
                considered = ci.reshape(L, K, order="F")
            else:
                considered = "All"
            dpr = LK / np.sum(np.abs(C) ** 2)                                     # :88
            C = C * np.sqrt(dpr)                                                  # :89
            power = np.sum(np.abs(C) ** 2, axis=1)                                # :98
            self.AuxiliaryToDataPowerOffset = power[aux].mean() / power[dat].mean()   # :100
            self.PostCodingChannelMatrix = np.nan
        elif Method == "Coding":                                                  # :106-210
            nD, nA = LK - 2 * P, 0
            self.AuxiliaryToDataPowerOffset = 0
            tmp, ci = _considered(D0, PilotMatrix, NrCanceledInterferersPerPilot, IM, tie_rtol)
            if np.sum(tmp.sum(axis=0) > 1):                                       # :116-118
                raise ValueError("Coding symbols must not overlap: The pilot-spacing is too small!")
            unc = np.flatnonzero(ci == 0)
            nU = len(unc)                                                         # :123
            C = np.zeros((LK, LK - P), dtype=complex)                             # :125
            C[pil, np.arange(P)] = np.sqrt(PilotToDataPowerOffset)                # :126
            C[unc, P + np.arange(nU)] = 1.0                                       # :127
            col_outer = P + nU                                                    # :129
            for ip in range(1, P + 1):                                            # :130-198
                pos = np.flatnonzero(ci == -ip)
                row = np.flatnonzero(ci == ip)
                interf = D0[np.ix_(row, pos)].reshape(-1)                         # :131
                interf = np.floor(np.abs(interf.imag * 1e10) + 0.5) * np.sign(interf.imag) / 1e10   # :133
                n = len(interf)                                                   # :135
                order = np.argsort(-np.abs(interf), kind="stable")                # :137
                abs_sorted = np.abs(interf)[order]
                isort = interf[order]                                             # :138
                uniq = np.unique(np.abs(isort))                                   # :140 (ascending)
                counts = np.array([np.sum(abs_sorted == u) for u in uniq])        # hist at exact centers
                C1 = np.zeros((n, n - 1))                                         # :142
                col = 0
                for u, nc in zip(uniq, counts):                                   # :144-166
                    sel = abs_sorted == u
                    it = isort[sel]
                    if np.log2(nc) % 1 == 0:                                      # :149 (incl. nc == 1)
                        ct = hadamard(int(nc)).astype(float) / it[:, None]        # :151
                        ct = ct[:, 1:]                                            # :152
                    elif nc > 1:                                                  # :156-161
                        e = np.eye(int(nc), int(nc) - 1)
                        ct = e / it[:, None] - np.roll(e, 1, axis=0) / it[:, None]
                    else:
                        continue
                    C1[np.ix_(np.flatnonzero(sel), col + np.arange(ct.shape[1]))] = ct
                    col += ct.shape[1]
                clusters = [(abs_sorted == u).astype(float) for u in uniq]       # :169
                for _ in range(len(uniq) - 1):                                    # :170-182
                    i1 = int(np.argmin([c.sum() for c in clusters]))
                    c1 = clusters.pop(i1)
                    i2 = int(np.argmin([c.sum() for c in clusters]))
                    c2 = clusters.pop(i2)
                    comb = [int(np.flatnonzero(c1)[0]), int(np.flatnonzero(c2)[0])]
                    clusters.append(c1 + c2)
                    C1[comb, col] = np.array([1.0, -1.0]) / isort[comb]           # :181
                    col += 1
                CG = np.zeros_like(C1)                                            # :185-191 Gram-Schmidt
                CG[:, 0] = C1[:, 0] / np.sqrt(C1[:, 0] @ C1[:, 0])
                for ig in range(1, n - 1):
                    v = C1[:, ig]
                    w = v - CG[:, :ig] @ (v @ CG[:, :ig])
                    CG[:, ig] = w / np.sqrt(w @ w)
                res = np.zeros_like(CG)                                           # :192-193
                res[order, :] = CG
                C[np.ix_(pos, col_outer + np.arange(n - 1))] = res                # :196
                col_outer += n - 1                                                # :197
            dpr = LK / np.sum(np.abs(C) ** 2)                                     # :200
            C = C * np.sqrt(dpr)                                                  # :201
            considered = ci.reshape(L, K, order="F")
            self.PostCodingChannelMatrix = np.abs(C.conj().T) ** 2                # :210
        else:
            raise ValueError("Method must be 'Auxiliary' or 'Coding'!")
        tmpm = D0[pil, :] @ C                                                     # :92 / :203
        dg = np.abs(np.diag(tmpm[:, :P])) ** 2
        self.SIR_dB = 10 * np.log10(dg / (np.sum(np.abs(tmpm) ** 2, axis=1) - dg))   # :95 / :206
        self.Method = Method
        self.PilotMatrix = PilotMatrix
        self.PrecodingMatrix = C
        self.NrDataSymbols = nD
        self.NrPilotSymbols = P
        self.NrAuxiliarySymbols = nA
        self.NrTransmittedSymbols = C.shape[0]
        self.PilotToDataPowerOffset = PilotToDataPowerOffset
        self.DataPowerReduction = dpr
        self.ConsideredInterferenceMatrix = considered
