"""Test helpers: feed a DeviceContext with the ORACLE's setup outputs, so that kernels are
compared with the oracle on identical inputs (G, Q, C, W, tables)."""
import numpy as np

import chest_b200


def mmse_csc_arrays(S, sc, key):
    """W_MMSE (K^2 P x n_snr) in MATLAB CSC form from the oracle's support-compressed W."""
    m = S["schemes"][sc]
    w = S["wf"][m["waveform"]]
    K2 = w["K"] ** 2
    jc, ir, val = [0], [], []
    for Wv in m[key]:
        p_idx, a_idx = np.nonzero(Wv.T)                      # ascending p, then ascending support index
        ir.append(w["sup"][a_idx].astype(np.int64) + K2 * p_idx.astype(np.int64))
        val.append(Wv.T[p_idx, a_idx])
        jc.append(jc[-1] + len(a_idx))
    return np.array(jc, dtype=np.int64), np.concatenate(ir), np.concatenate(val)


def context_from_oracle(S, max_batch=32, device=0):
    cfg = S["cfg"]
    ctx = chest_b200.DeviceContext(device)
    chan = S["chan"]
    ctx.set_channel(S["N"], chan.Implementation["PowerDelayProfileNormalized"], S["fD"], chan.PHY["dt"],
                    cfg.Paths, cfg.DopplerModel)
    for name, w in S["wf"].items():
        ctx.set_waveform(name, w["G"], w["Q"])
    ctx.set_constellation("PAM", S["PAM"].SymbolMapping, S["PAM"].BitMapping)
    ctx.set_constellation("QAM", S["QAM"].SymbolMapping, S["QAM"].BitMapping)
    ctx.set_snr(S["Pn"])
    detect = {"aux": "select_real", "cod": "despread_real", "ofdm": "select_complex"}
    for sc, m in S["schemes"].items():
        w = S["wf"][m["waveform"]]
        ctx.set_scheme(sc, m["waveform"], m["C"], w["pil"], m["data_idx"], m["kappa"], m["dpr"], detect[sc],
                       m["const"], m["considered_bits"])
        for variant, key in ((0, "W"), (1, "W_noInt")):
            ctx.set_mmse_arrays(sc, variant, *mmse_csc_arrays(S, sc, key))
    ctx.finalize(max_batch)
    return ctx


def err_from_oracle(out, n_iter):
    """oracle ds_realization output -> err[snr, it, scheme, csi, edge] like the library's layout."""
    sid = {"aux": 0, "cod": 1, "ofdm": 2}
    nS = next(iter(out["err"].values()))["est"]["all"].shape[0]
    e = np.zeros((nS, n_iter + 1, 3, 2, 2), dtype=np.uint32)
    for sc, d in out["err"].items():
        for ci, csi in enumerate(("est", "perfect")):
            for ei, edge in enumerate(("all", "noedge")):
                e[:, :, sid[sc], ci, ei] = d[csi][edge][:, :n_iter + 1]
    return e
