"""Host mirror of the reference's +ChannelEstimation package.

`ImaginaryInterferenceCancellationAtPilotPosition` (IIC.m): builds the FBMC precoding matrix
(auxiliary symbols or data spreading).  One-time host-side setup, as in the reference; its
`PrecodingMatrix` is consumed by the device kernels through chest_set_scheme.
`PilotSymbolAidedChannelEstimation` (PSACE.m): pilot patterns and matrix-form interpolators."""
import numpy as np


def _sylvester_hadamard(n):
    H = np.ones((1, 1))
    while H.shape[0] < n:
        H = np.block([[H, H], [H, -H]])
    return H


class ImaginaryInterferenceCancellationAtPilotPosition:
    """ChannelEstimation.ImaginaryInterferenceCancellationAtPilotPosition(Method, PilotMatrix,
    FBMCMatrix, NrCanceledInterferersPerPilot, PilotToDataPowerOffset)  (IIC.m:37-229)."""

    # Relative width of a tie group around the selection threshold (see _tags).  0 = the reference's literal >=.
    TieTolerance = 1e-9

    def __init__(self, Method, PilotMatrix, FBMCMatrix, NrCanceledInterferersPerPilot, PilotToDataPowerOffset,
                 TieTolerance=None):
        if TieTolerance is not None:
            self.TieTolerance = float(TieTolerance)
        pm2 = np.asarray(PilotMatrix)
        D0 = np.asarray(FBMCMatrix)
        L, K = pm2.shape
        LK = L * K
        pm = pm2.reshape(-1, order="F")
        pil = np.flatnonzero(pm == 1)
        P = len(pil)
        self.Method, self.PilotMatrix = Method, pm2
        self.NrPilotSymbols = P
        self.PilotToDataPowerOffset = PilotToDataPowerOffset
        n_cancel = int(NrCanceledInterferersPerPilot)
        if Method == "Auxiliary":
            dat, aux = np.flatnonzero(pm == 0), np.flatnonzero(pm == -1)
            nD, nA = len(dat), len(aux)
            inv = np.linalg.pinv(D0[np.ix_(pil, aux)])                                   # IIC.m:60
            C = np.zeros((LK, LK - nA), dtype=np.complex128)
            C[np.ix_(aux, np.arange(P))] = inv @ (np.eye(P) - D0[np.ix_(pil, pil)])     # IIC.m:61,66
            C[np.ix_(aux, P + np.arange(nD))] = -inv @ D0[np.ix_(pil, dat)]             # IIC.m:62,67
            C[pil, np.arange(P)] = np.sqrt(PilotToDataPowerOffset)                       # IIC.m:68
            C[dat, P + np.arange(nD)] = 1.0                                              # IIC.m:69
            if n_cancel > 0:                                                             # IIC.m:71-82
                tags = self._tags(D0, pm2, pil, n_cancel, self.TieTolerance)[1]
                keep = np.concatenate([tags[pil], tags[dat]]) != 0
                C[np.ix_(aux, np.flatnonzero(~keep))] = 0
                self.ConsideredInterferenceMatrix = tags.reshape(L, K, order="F")
            else:
                self.ConsideredInterferenceMatrix = "All"
            self.NrDataSymbols, self.NrAuxiliarySymbols = nD, nA
            self.PostCodingChannelMatrix = np.nan
        elif Method == "Coding":
            mask, tags = self._tags(D0, pm2, pil, n_cancel, self.TieTolerance)
            if np.any(mask.sum(axis=0) > 1):                                             # IIC.m:116-118
                raise ValueError("Coding symbols must not overlap: The pilot-spacing is too small!")
            free = np.flatnonzero(tags == 0)
            C = np.zeros((LK, LK - P), dtype=np.complex128)
            C[pil, np.arange(P)] = np.sqrt(PilotToDataPowerOffset)                       # IIC.m:126
            C[free, P + np.arange(len(free))] = 1.0                                      # IIC.m:127
            col0 = P + len(free)
            for ip in range(1, P + 1):                                                   # IIC.m:130-198
                pos = np.flatnonzero(tags == -ip)
                code = self._spreading_code(D0[pil[ip - 1], pos])
                C[np.ix_(pos, col0 + np.arange(code.shape[1]))] = code
                col0 += code.shape[1]
            self.NrDataSymbols, self.NrAuxiliarySymbols = LK - 2 * P, 0
            self.ConsideredInterferenceMatrix = tags.reshape(L, K, order="F")
        else:
            raise ValueError("Method must be  'Auxiliary' or 'Coding'!")
        self.DataPowerReduction = LK / np.sum(np.abs(C) ** 2)                            # IIC.m:88,200
        C *= np.sqrt(self.DataPowerReduction)
        if Method == "Auxiliary":
            power = np.sum(np.abs(C) ** 2, axis=1)
            self.AuxiliaryToDataPowerOffset = power[pm == -1].mean() / power[pm == 0].mean()   # IIC.m:100
        else:
            self.AuxiliaryToDataPowerOffset = 0
            self.PostCodingChannelMatrix = np.abs(C.conj().T) ** 2                       # IIC.m:210
        T = D0[pil, :] @ C                                                               # IIC.m:92-96 / 203-207
        sig = np.abs(T[np.arange(P), np.arange(P)]) ** 2
        self.SIR_dB = 10 * np.log10(sig / (np.sum(np.abs(T) ** 2, axis=1) - sig))
        self.PrecodingMatrix = C
        self.NrTransmittedSymbols = LK

    @staticmethod
    def _tags(D0, pm2, pil, n_cancel, tie=0.0):
        """IIC.m:47-51,72-76,113-122: positions whose interference weight towards pilot p is among
        the n_cancel largest weights of the interference pattern get tag -p, pilots get +p.
        The reference compares nominally equal weights with a plain >=; when the threshold falls inside
        a group of weights that are equal in exact arithmetic (FFT size = subcarrier count: offsets +l
        and L-l alias), which members pass depends on their last bits, i.e. on the FFT library.  `tie`
        > 0 takes the exact-arithmetic reading instead: the whole group passes."""
        L, K = pm2.shape
        LK = L * K
        corner = lambda c: np.abs(D0[:, c]).reshape(L, K, order="F")
        i11, iE1, i1E, iEE = corner(0), corner(L - 1), corner(LK - L), corner(LK - 1)
        pattern = np.hstack([np.vstack([iEE, i1E[1:, :]]), np.vstack([iE1[:, 1:], i11[1:, 1:]])])
        thr = np.sort(pattern.reshape(-1))[::-1][n_cancel]
        mask = np.abs(D0[pil, :]) >= thr * (1.0 - tie)
        tags = -(mask * np.arange(1, len(pil) + 1)[:, None]).sum(axis=0).astype(np.int64)
        tags[pil] = np.arange(1, len(pil) + 1)
        return mask, tags

    @staticmethod
    def _spreading_code(row):
        """IIC.m:131-193: orthonormal code over the interferers of one pilot whose weighted sum
        (weights = imaginary interference) vanishes: Hadamard codes inside clusters of equal
        |interference|, pairwise links between clusters, then Gram-Schmidt."""
        w = np.imag(row)
        w = np.floor(np.abs(w) * 1e10 + 0.5) * np.sign(w) / 1e10                         # IIC.m:133
        n = len(w)
        order = np.argsort(-np.abs(w), kind="stable")                                    # IIC.m:137
        ws = w[order]
        mags = np.abs(ws)
        levels = np.unique(mags)                                                         # ascending, IIC.m:140
        B = np.zeros((n, n - 1))
        col = 0
        members = []
        for u in levels:                                                                 # IIC.m:144-166
            idx = np.flatnonzero(mags == u)
            members.append(idx)
            m = len(idx)
            if (m & (m - 1)) == 0:                                                       # power of two
                blk = (_sylvester_hadamard(m) / ws[idx][:, None])[:, 1:]
            elif m > 1:
                e = np.eye(m, m - 1)
                blk = (e - np.roll(e, 1, axis=0)) / ws[idx][:, None]
            else:
                continue
            B[np.ix_(idx, col + np.arange(blk.shape[1]))] = blk
            col += blk.shape[1]
        groups = [np.isin(np.arange(n), idx).astype(float) for idx in members]           # IIC.m:169-182
        for _ in range(len(levels) - 1):
            a = groups.pop(int(np.argmin([g.sum() for g in groups])))
            b = groups.pop(int(np.argmin([g.sum() for g in groups])))
            ia, ib = int(np.flatnonzero(a)[0]), int(np.flatnonzero(b)[0])
            B[[ia, ib], col] = np.array([1.0, -1.0]) / ws[[ia, ib]]
            col += 1
            groups.append(a + b)
        Qm = np.zeros_like(B)                                                            # IIC.m:185-191
        for c in range(n - 1):
            v = B[:, c] - Qm[:, :c] @ (B[:, c] @ Qm[:, :c])
            Qm[:, c] = v / np.sqrt(v @ v)
        out = np.zeros_like(Qm)
        out[order, :] = Qm                                                               # IIC.m:192-193
        return out


class PilotSymbolAidedChannelEstimation:
    """ChannelEstimation.PilotSymbolAidedChannelEstimation(PilotPattern, PatternParameters,
    InterpolationMethod[, BlockLengths])  (PSACE.m:33-113).  Patterns: 'Rectangular', 'Diamond',
    'Custom'.  Interpolators in matrix form: 'FullAverage', 'MovingBlockAverage', and 'linear' /
    'nearest' through scipy (the reference's scatteredInterpolant is closed-source MATLAB; its
    triangulation and extrapolation rules are not reproducible bit for bit -- parity unpinned)."""

    def __init__(self, PilotPattern, PatternParameters, InterpolationMethod, BlockLengths=None):
        self.PilotPattern, self.InterpolationMethod = PilotPattern, InterpolationMethod
        rnd = lambda x: int(np.floor(x + 0.5))
        if PilotPattern in ("Rectangular", "Diamond"):
            prm = np.asarray(PatternParameters, dtype=float)
            nL, dF, nK, dT = int(prm[0, 0]), prm[0, 1], int(prm[1, 0]), prm[1, 1]
            self.PilotSpacingFrequency, self.PilotSpacingTime = dF, dT
            pm = np.zeros((nL, nK))
            if PilotPattern == "Rectangular":                                            # PSACE.m:47-48
                f0 = rnd(((nL - 1) % dF) / 2)
                t0 = rnd(rnd(((nK - 1) % dT) / 2))
                pm[np.ix_(np.arange(f0, nL, int(dF)), np.arange(t0, nK, int(dT)))] = 1
            else:                                                                        # PSACE.m:55-62
                def last(start, step, n):                                                # max(start:step:n), 1-based
                    return start + np.floor((n - start) / step) * step if start <= n else -np.inf
                fmax = max(last(1, 2 * dF, nL), last(1 + dF / 2, 2 * dF, nL), last(1 + dF, 2 * dF, nL),
                           last(1 + 3 * dF / 2, 2 * dF, nL))
                tmax = max(last(1, 2 * dT, nK), last(1 + dT, 2 * dT, nK))
                fs = int(np.floor((nL - fmax) / 2)) + 1
                ts = int(np.floor((nK - tmax) / 2)) + 1
                rows = lambda off: np.arange(fs + rnd(off) - 1, nL, int(2 * dF))
                cols = lambda off: np.arange(rnd(ts + off) - 1, nK, int(2 * dT))
                pm[np.ix_(rows(0), cols(0))] = 1
                pm[np.ix_(rows(dF / 2), cols(dT))] = 1
                pm[np.ix_(rows(dF), cols(0))] = 1
                pm[np.ix_(rows(3 * dF / 2), cols(dT))] = 1
        elif PilotPattern == "Custom":
            self.PilotSpacingFrequency = self.PilotSpacingTime = np.nan
            pm = np.asarray(PatternParameters, dtype=float)
        else:
            raise ValueError("Pilot pattern is not supported! Chose Rectangular Diamond or Custom")
        self.PilotMatrix = pm
        self.NrPilotSymbols = int(np.sum(pm))
        self.InterpolationProperties = {}
        if InterpolationMethod == "MovingBlockAverage":                                  # PSACE.m:78-109
            bF, bT = int(BlockLengths[0]), int(BlockLengths[1])
            nL, nK = pm.shape
            num = np.zeros(pm.shape, dtype=int)
            num[pm.astype(bool)] = 0
            flat_idx = np.flatnonzero(pm.reshape(-1, order="F"))
            numbered = -np.ones(pm.size, dtype=int)
            numbered[flat_idx] = np.arange(len(flat_idx))
            numbered = numbered.reshape(pm.shape, order="F")
            M = np.zeros((pm.size, self.NrPilotSymbols))
            for pos in range(pm.size):
                f, t = pos % nL, pos // nL
                blk = numbered[max(f - bF, 0):f + bF + 1, max(t - bT, 0):t + bT + 1]
                sel = blk[blk >= 0]
                M[pos, sel] = 1.0 / len(sel)
            self.InterpolationProperties["InterpolationMatrix"] = M
        elif InterpolationMethod == "MMSE":
            raise NotImplementedError("Needs to be implemented")                         # PSACE.m:110-111
        elif InterpolationMethod not in ("linear", "nearest", "natural", "FullAverage"):
            raise ValueError("Interpolation method not implemented")

    def ChannelInterpolation(self, LSChannelEstimatesAtPilotPosition):
        """PSACE.m:115-133: interpolate P pilot estimates over the L x K grid."""
        v = np.asarray(LSChannelEstimatesAtPilotPosition).reshape(-1)
        pm = self.PilotMatrix
        if self.InterpolationMethod == "FullAverage":
            return np.ones(pm.shape) * np.mean(v)
        if self.InterpolationMethod == "MovingBlockAverage":
            return (self.InterpolationProperties["InterpolationMatrix"] @ v).reshape(pm.shape, order="F")
        from scipy.interpolate import LinearNDInterpolator, NearestNDInterpolator
        fpos, tpos = np.nonzero(pm.T)[1], np.nonzero(pm.T)[0]                             # column-major find()
        pts = np.column_stack([fpos, tpos]).astype(float)
        ff, tt = np.meshgrid(np.arange(pm.shape[0]), np.arange(pm.shape[1]), indexing="ij")
        near = NearestNDInterpolator(pts, v)(ff, tt)
        if self.InterpolationMethod == "nearest":
            return near
        lin = LinearNDInterpolator(pts, v)(ff, tt)
        return np.where(np.isnan(lin), near, lin)        # outside the convex hull: nearest (MATLAB extrapolates linearly)

    def GetInterpolationMatrix(self):
        """PSACE.m:171-184."""
        P = self.NrPilotSymbols
        M = np.zeros((self.PilotMatrix.size, P), dtype=complex)
        for i in range(P):
            e = np.zeros(P)
            e[i] = 1
            M[:, i] = self.ChannelInterpolation(e).reshape(-1, order="F")
        return M

    def GetAuxiliaryMatrix(self, NrAxuiliarySymbols):
        """PSACE.m:137-169."""
        if NrAxuiliarySymbols not in (1, 2, 3, 4):
            raise ValueError("Only 1,2,3,4 auxiliary symbols per pilot are supported")
        A = self.PilotMatrix.copy()
        ls, ks = np.nonzero(self.PilotMatrix)
        for l, k in zip(ls, ks):
            A[l, k + 1] = -1
            if NrAxuiliarySymbols >= 2:
                A[l, k - 1] = -1
            if NrAxuiliarySymbols >= 3:
                A[l + 1, k] = -1
            if NrAxuiliarySymbols >= 4:
                A[l - 1, k] = -1
        return A
