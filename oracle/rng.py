"""Oracle restatement of the product's counter-based generator (Philox4x32-10), so that seeded
runs can be reproduced on the CPU.  The reference itself is unseeded (DS.m:12).  Streams and
counter layout are documented in DESIGN.md "Random numbers".
Test infrastructure only -- see oracle/__init__.py."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
RS_DOPPLER, RS_PHASE, RS_BITS0, RS_PILOT0, RS_NOISE = 0, 1, 2, 5, 7
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def _u53(lo, hi):
    w = (hi << np.uint64(32)) | lo
    return (w >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) + (0.5 / 9007199254740992.0)


def _ctr(seed, rep, stream, n_blocks):
    r = int(rep) & 0xFFFFFFFFFFFFFFFF
    return philox4x32_10(np.arange(n_blocks, dtype=np.uint64), r & 0xFFFFFFFF, stream, r >> 32,
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def uniforms(seed, rep, stream, n):
    v = _ctr(seed, rep, stream, (n + 1) // 2)
    out = np.empty(2 * len(v[0]))
    out[0::2], out[1::2] = _u53(v[0], v[1]), _u53(v[2], v[3])
    return out[:n]


def bits(seed, rep, stream, n):
    v = _ctr(seed, rep, stream, (n + 127) // 128)
    words = np.stack(v, axis=1).astype(np.uint64)                 # (blocks, 4)
    t = np.arange(32, dtype=np.uint64)
    b = ((words[:, :, None] >> t[None, None, :]) & np.uint64(1)).reshape(-1)
    return b[:n].astype(np.uint8)


def indices(seed, rep, stream, n, order):
    u = uniforms(seed, rep, stream, n)
    return np.minimum(np.floor(u * order).astype(np.int32), order - 1)


def normals(seed, rep, snr, n):
    v = _ctr(seed, rep, RS_NOISE | (snr << 8), n)
    u1, u2 = _u53(v[0], v[1]), _u53(v[2], v[3])
    rad = np.sqrt(-2.0 * np.log(u1))
    return rad * np.cos(2 * np.pi * u2) + 1j * rad * np.sin(2 * np.pi * u2)


def draws_for(S, seed, rep):
    """The draws the device generator produces for realization `rep` (same dict as ds.new_draws)."""
    cfg = S["cfg"]
    T = len(S["chan"].Implementation["IndexDelayTaps"])
    n = T * cfg.Paths
    d = dict(doppler_u=uniforms(seed, rep, RS_DOPPLER, n).reshape(T, cfg.Paths, order="F"),
             phase_u=uniforms(seed, rep, RS_PHASE, n).reshape(T, cfg.Paths, order="F"))
    for sid, sc in enumerate(("aux", "cod", "ofdm")):
        if sc in S["schemes"]:
            m = S["schemes"][sc]
            d["bits_" + sc] = bits(seed, rep, RS_BITS0 + sid, m["nD"] * m["nbits"])
    d["pil_idx_fbmc"] = indices(seed, rep, RS_PILOT0 + 0, S["P"], S["PAM"].ModulationOrder)
    d["pil_idx_ofdm"] = indices(seed, rep, RS_PILOT0 + 1, S["P"], S["QAM"].ModulationOrder)
    d["noise"] = np.stack([normals(seed, rep, s, S["N"]) for s in range(len(cfg.M_SNR_dB))])
    return d
