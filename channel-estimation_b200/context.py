"""Thin NumPy-facing wrapper of the C ABI (one `DeviceContext` == one `chest_create` handle).

All arithmetic happens in the CUDA library; this file only marshals arrays (column-major,
complex as interleaved doubles, 0-based indices) and raises on any non-zero status."""
import ctypes as C
import numpy as np
import scipy.sparse as sp

from . import _lib

SCHEME_ID = {"aux": 0, "cod": 1, "ofdm": 2}
SCHEME_NAME = {0: "aux", 1: "cod", 2: "ofdm"}
WF_ID = {"F": 0, "O": 1}
DETECT = {"select_real": 0, "despread_real": 1, "select_complex": 2}
CONST_ID = {"PAM": 0, "QAM": 1}


class ChestError(RuntimeError):
    pass


def _c(a):
    """complex array -> contiguous complex128 (memory = interleaved doubles)"""
    return np.ascontiguousarray(a, dtype=np.complex128)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class DeviceContext:
    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_uint64(0)
        self._h = None
        self._check(self.lib.chest_create(device, C.byref(h)))
        self._h = h
        self.device = device
        self.schemes = {}
        self.n_snr = 0
        self.N = 0
        self.K = {}

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise ChestError("chest_b200 error %d: %s" % (rc, self.lib.chest_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.lib.chest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ setup
    def set_channel(self, n_samples, pdp_normalized, max_doppler, dt, paths, model="Jakes"):
        pdp = np.ascontiguousarray(pdp_normalized, dtype=np.float64)
        self._check(self.lib.chest_set_channel(self._h, n_samples, len(pdp), _ptr(pdp), max_doppler, dt, paths,
                                               {"Jakes": 0, "Uniform": 1, "Discrete-Jakes": 2, "Discrete-Uniform": 3}[model]))
        self.N, self.Lt = n_samples, len(pdp)
        self.T = int(np.count_nonzero(pdp))
        self.paths = paths
        ns, fd = C.c_int(0), C.c_double(0)
        self._check(self.lib.chest_channel_info(self._h, C.byref(ns), C.byref(fd), None))
        self.n_doppler_shifts, self.max_doppler = ns.value, fd.value

    def set_waveform(self, wf, G, Q):
        G, Q = np.asfortranarray(G, dtype=np.complex128), np.asfortranarray(Q, dtype=np.complex128)
        assert G.shape == Q.shape
        self._check(self.lib.chest_set_waveform(self._h, WF_ID[wf], G.shape[0], G.shape[1], _ptr(G), _ptr(Q)))
        self.N = G.shape[0]
        self.K[wf] = G.shape[1]

    def set_constellation(self, which, symbol_mapping, bit_mapping):
        s = _c(symbol_mapping)
        b = np.asfortranarray(bit_mapping, dtype=np.uint8)
        self._check(self.lib.chest_set_constellation(self._h, CONST_ID[which], len(s), _ptr(s), _ptr(b)))

    def set_scheme(self, name, wf, C_mat, pilot_pos, data_pos, kappa, dpr, detect, constellation, considered_bits):
        Cs = sp.csc_matrix(C_mat).astype(np.complex128)
        Cs.sort_indices()
        jc = Cs.indptr.astype(np.int64)
        ir = Cs.indices.astype(np.int32)
        val = _c(Cs.data)
        pp = np.ascontiguousarray(pilot_pos, dtype=np.int32)
        dp = np.ascontiguousarray(data_pos, dtype=np.int32) if data_pos is not None else None
        cb = np.ascontiguousarray(considered_bits, dtype=np.uint8)
        n_data = Cs.shape[1] - len(pp)
        self._check(self.lib.chest_set_scheme(self._h, SCHEME_ID[name], WF_ID[wf], Cs.shape[1], len(pp), n_data,
                                              _ptr(jc), _ptr(ir), _ptr(val), _ptr(pp), _ptr(dp), kappa, dpr,
                                              DETECT[detect], CONST_ID[constellation], _ptr(cb)))
        self.schemes[name] = dict(wf=wf, K=Cs.shape[0], P=len(pp), n_data=n_data, n_bits=len(cb))

    def set_snr(self, pn_time):
        pn = np.ascontiguousarray(pn_time, dtype=np.float64)
        self._check(self.lib.chest_set_snr(self._h, len(pn), _ptr(pn)))
        self.n_snr = len(pn)

    def set_mmse(self, name, variant, W_csc):
        """W_csc: scipy CSC of shape (K^2 P, n_snr) -- W_MMSE_* as DS.m:279-313 stores it."""
        W = sp.csc_matrix(W_csc)
        W.sort_indices()
        jc = W.indptr.astype(np.int64)
        ir = W.indices.astype(np.int64)
        val = _c(W.data)
        self._check(self.lib.chest_set_mmse(self._h, SCHEME_ID[name], variant, W.shape[1], _ptr(jc), _ptr(ir), _ptr(val)))

    def set_mmse_arrays(self, name, variant, jc, ir, val):
        jc = np.ascontiguousarray(jc, dtype=np.int64)
        ir = np.ascontiguousarray(ir, dtype=np.int64)
        val = _c(val)
        self._check(self.lib.chest_set_mmse(self._h, SCHEME_ID[name], variant, len(jc) - 1, _ptr(jc), _ptr(ir), _ptr(val)))

    def setup_correlations(self, wf, pilot_pos, time_correlation, zero_threshold):
        """DS.m:208-268 on the device: returns (R_hP, number of (i,j) with a non-zero row of R_Dij_hP); R_Dij_hP stays
        on the device for build_mmse."""
        pp = np.ascontiguousarray(pilot_pos, dtype=np.int32)
        rt = np.ascontiguousarray(time_correlation, dtype=np.float64)
        assert rt.shape == (2 * self.N - 1,)
        R = np.zeros((len(pp), len(pp)), dtype=np.complex128)
        n = C.c_int64(0)
        self._check(self.lib.chest_setup_correlations(self._h, WF_ID[wf], len(pp), _ptr(pp), _ptr(rt), zero_threshold,
                                                      _ptr(R), C.byref(n)))
        return R.T.copy(), int(n.value)          # library writes column-major

    def build_mmse(self, name, variant, R_inv, zero_threshold):
        """W = R_Dij_hP * R_inv[snr] for every SNR point, thresholded, written as device tiles (DS.m:283-313).
        R_inv: (n_snr, P, P) complex."""
        Ri = _c(np.transpose(np.asarray(R_inv), (0, 2, 1)))      # per SNR column-major P x P
        self._check(self.lib.chest_build_mmse(self._h, SCHEME_ID[name], variant, Ri.shape[0], _ptr(Ri), zero_threshold))

    def release_setup(self):
        self._check(self.lib.chest_release_setup(self._h))

    def finalize(self, max_batch):
        self._check(self.lib.chest_finalize(self._h, max_batch))
        self.max_batch = max_batch

    # ------------------------------------------------------------------ tier 1
    def new_realization(self, doppler_u, phase_u):
        du = np.ascontiguousarray(doppler_u, dtype=np.float64)
        pu = np.ascontiguousarray(phase_u, dtype=np.float64)
        batch = du.size // (self.T * self.paths)
        self._check(self.lib.chest_new_realization(self._h, batch, _ptr(du), _ptr(pu)))

    def new_realization_gauss(self, gauss):
        """'Discrete-*' Doppler models: gauss (batch, 2 n_shift + 1, T) complex standard normals (FF.m:208-209)."""
        g = _c(gauss)
        assert g.shape[1:] == (2 * self.n_doppler_shifts + 1, self.T)
        self._check(self.lib.chest_new_realization_gauss(self._h, g.shape[0], _ptr(g)))

    def new_realization_seeded(self, batch, seed, first_rep=0):
        self._check(self.lib.chest_new_realization_seeded(self._h, batch, seed, first_rep))

    def set_impulse_response(self, h):
        """h: (batch, N, Lt) complex."""
        h = np.asarray(h)
        hb = _c(np.transpose(h, (0, 2, 1)))          # per realization column-major N x Lt
        self._check(self.lib.chest_set_impulse_response(self._h, h.shape[0], _ptr(hb)))

    def impulse_response(self, b=0):
        out = np.zeros((self.Lt, self.N), dtype=np.complex128)
        self._check(self.lib.chest_get_impulse_response(self._h, b, _ptr(out)))
        return out.T.copy()

    def convolution_matrix(self, b=0):
        nnz = C.c_int64(0)
        self._check(self.lib.chest_get_convolution_csc(self._h, b, C.byref(nnz), None, None, None))
        jc = np.zeros(self.N + 1, dtype=np.int64)
        ir = np.zeros(nnz.value, dtype=np.int32)
        val = np.zeros(nnz.value, dtype=np.complex128)
        self._check(self.lib.chest_get_convolution_csc(self._h, b, C.byref(nnz), _ptr(jc), _ptr(ir), _ptr(val)))
        return sp.csc_matrix((val, ir, jc), shape=(self.N, self.N))

    def convolve(self, s, b=0):
        s2 = np.atleast_2d(np.asarray(s).T) if np.ndim(s) == 1 else np.asarray(s).T
        s2 = _c(s2)
        r = np.zeros_like(s2)
        self._check(self.lib.chest_convolve(self._h, b, _ptr(s2), s2.shape[0], _ptr(r)))
        return r[0] if np.ndim(s) == 1 else r.T.copy()

    def transmission_matrix(self, wf, b=0):
        K = self.K[wf]
        D = np.zeros((K, K), dtype=np.complex128)
        h = np.zeros(K, dtype=np.complex128)
        self._check(self.lib.chest_transmission_matrix(self._h, b, WF_ID[wf], _ptr(D), _ptr(h)))
        return D.T.copy(), h        # library writes column-major

    def transmission_matrix_batch(self, wf, n_rep, h_out_ptr=None):
        """D = Q' H_b G for realizations 0..n_rep-1 on the device (left there).  Returns (ms of k_apply_hg, ms of k_gemm_d,
        support-aware flops per realization); h_out_ptr: host address receiving diag(D) (n_rep x K complex)."""
        ms = (C.c_float * 2)()
        fl = C.c_double(0)
        self._check(self.lib.chest_transmission_matrix_batch(self._h, n_rep, WF_ID[wf], ms, C.byref(fl), C.c_void_p(h_out_ptr)))
        return float(ms[0]), float(ms[1]), fl.value

    def transmission_matrix_entries(self, wf, b, rows, cols):
        rows = np.ascontiguousarray(rows, dtype=np.int32); cols = np.ascontiguousarray(cols, dtype=np.int32)
        out = np.zeros(len(rows), dtype=np.complex128)
        self._check(self.lib.chest_transmission_matrix_entries(self._h, b, WF_ID[wf], len(rows), rows.ctypes.data, cols.ctypes.data, _ptr(out)))
        return out

    def modulate(self, wf, x):
        x2 = _c(np.atleast_2d(np.asarray(x).T) if np.ndim(x) == 1 else np.asarray(x).T)
        s = np.zeros((x2.shape[0], self.N), dtype=np.complex128)
        self._check(self.lib.chest_modulate(self._h, WF_ID[wf], _ptr(x2), x2.shape[0], _ptr(s)))
        return s[0] if np.ndim(x) == 1 else s.T.copy()

    def demodulate(self, wf, r):
        r2 = _c(np.atleast_2d(np.asarray(r).T) if np.ndim(r) == 1 else np.asarray(r).T)
        y = np.zeros((r2.shape[0], self.K[wf]), dtype=np.complex128)
        self._check(self.lib.chest_demodulate(self._h, WF_ID[wf], _ptr(r2), r2.shape[0], _ptr(y)))
        return y[0] if np.ndim(r) == 1 else y.T.copy()

    # ------------------------------------------------------------------ FFT modem
    def set_modem(self, wf, kind, L, Ksym, nfft, bins, time_spacing, O=1, cp=0, zero_guard=0, prototype_filter=None,
                  phase_shift=None, normalization=1.0, subcarrier_spacing=1.0):
        """kind: 'fbmc' (polyphase, FBMC.m:255-302) or 'ofdm' (OFDM.m:153-181)."""
        b = np.ascontiguousarray(bins, dtype=np.int32)
        pf = np.ascontiguousarray(prototype_filter, dtype=np.float64) if prototype_filter is not None else None
        ph = _c(np.asfortranarray(phase_shift).reshape(-1, order="F")) if phase_shift is not None else None
        self._check(self.lib.chest_set_modem(self._h, WF_ID[wf], {"fbmc": 0, "ofdm": 1}[kind], L, Ksym, nfft, _ptr(b),
                                             time_spacing, O, cp, zero_guard, _ptr(pf), _ptr(ph), normalization,
                                             subcarrier_spacing))
        n = O * nfft + (Ksym - 1) * time_spacing if kind == "fbmc" else 2 * zero_guard + Ksym * time_spacing
        self._modem = getattr(self, "_modem", {})
        self._modem[wf] = (L * Ksym, n)

    def modulate_fft(self, wf, x):
        """x: (L*K,) or (L*K, n_cols) -> (N,) or (N, n_cols)."""
        lk, n = self._modem[wf]
        x2 = _c(np.atleast_2d(np.asarray(x).T) if np.ndim(x) == 1 else np.asarray(x).T)
        s = np.zeros((x2.shape[0], n), dtype=np.complex128)
        self._check(self.lib.chest_modulate_fft(self._h, WF_ID[wf], _ptr(x2), x2.shape[0], _ptr(s)))
        return s[0] if np.ndim(x) == 1 else s.T.copy()

    def demodulate_fft(self, wf, r):
        lk, n = self._modem[wf]
        r2 = _c(np.atleast_2d(np.asarray(r).T) if np.ndim(r) == 1 else np.asarray(r).T)
        y = np.zeros((r2.shape[0], lk), dtype=np.complex128)
        self._check(self.lib.chest_demodulate_fft(self._h, WF_ID[wf], _ptr(r2), r2.shape[0], _ptr(y)))
        return y[0] if np.ndim(r) == 1 else y.T.copy()

    # ------------------------------------------------------------------ SimpleVersion_DoublyFlat.m chain
    def set_interpolation(self, name, M):
        """M: K x P interpolation matrix (PSACE.m:171-184)."""
        Mf = _c(np.asfortranarray(np.asarray(M, dtype=np.complex128)).reshape(-1, order="F"))
        self._check(self.lib.chest_set_interpolation(self._h, SCHEME_ID[name], _ptr(Mf)))

    def sv_run_batch(self, pn_time, draws=None, seed=0, first_body=0):
        """SV.m:89-176 for len(pn_time) bodies.  draws: list of per-body dicts (oracle.sv.sv_new_draws layout) or None
        (device generator).  Returns err[body, 5] (aux, cod, FBMC perfect, OFDM, OFDM perfect)."""
        pn = np.ascontiguousarray(pn_time, dtype=np.float64)
        n = len(pn)
        err = np.zeros((n, 5), dtype=np.uint32)
        dptr, keep = None, {}
        if draws is not None:
            st = _lib.ChestSvDraws()
            for name, sid in SCHEME_ID.items():
                keep["b" + name] = np.ascontiguousarray(np.stack([d["bits_" + name] for d in draws]), dtype=np.uint8)
                st.bits[sid] = keep["b" + name].ctypes.data
            for key, wid in (("pil_idx_fbmc", 0), ("pil_idx_ofdm", 1)):
                keep[key] = np.ascontiguousarray(np.stack([d[key] for d in draws]), dtype=np.int32)
                st.pilot_idx[wid] = keep[key].ctypes.data
            keep["h"] = _c(np.array([d["h"] for d in draws]))
            st.h = _ptr(keep["h"])
            for key, wid in (("noise_fbmc", 0), ("noise_ofdm", 1)):
                keep[key] = _c(np.stack([d[key] for d in draws]))
                st.noise[wid] = keep[key].ctypes.data
            st.on_device = 0
            dptr = C.byref(st)
        self._check(self.lib.chest_sv_run_batch(self._h, n, _ptr(pn), dptr, seed, first_body, _ptr(err)))
        return err

    def estimate(self, name, variant, i_snr, hP, want_D=True):
        K = self.schemes[name]["K"]
        hP = _c(hP)
        D = np.zeros((K, K), dtype=np.complex128) if want_D else None
        hd = np.zeros(K, dtype=np.complex128)
        self._check(self.lib.chest_estimate(self._h, SCHEME_ID[name], variant, i_snr, _ptr(hP), _ptr(D), _ptr(hd)))
        return (D.T.copy() if want_D else None), hd

    # ------------------------------------------------------------------ tier 2
    def pack_draws(self, draws_list):
        """list of per-realization dicts (doppler_u, phase_u, bits_<scheme>, pil_idx_fbmc, pil_idx_ofdm,
        noise) -> (ChestDraws struct, keep-alive arrays)."""
        keep = {}
        st = _lib.ChestDraws()
        if getattr(self, "n_doppler_shifts", 0) > 0:                 # 'Discrete-*': complex normals instead of uniforms
            keep["cg"] = _c(np.stack([d["gauss"] for d in draws_list]))
            st.channel_gauss = _ptr(keep["cg"])
        else:
            keep["du"] = np.ascontiguousarray(np.stack([d["doppler_u"].reshape(-1, order="F") for d in draws_list]))
            keep["pu"] = np.ascontiguousarray(np.stack([d["phase_u"].reshape(-1, order="F") for d in draws_list]))
            st.doppler_u, st.phase_u = _ptr(keep["du"]), _ptr(keep["pu"])
        keep["noise"] = _c(np.stack([d["noise"] for d in draws_list]))
        st.noise = _ptr(keep["noise"])
        for name, sid in SCHEME_ID.items():
            if name in self.schemes:
                keep["b" + name] = np.ascontiguousarray(np.stack([d["bits_" + name] for d in draws_list]), dtype=np.uint8)
                st.bits[sid] = keep["b" + name].ctypes.data
        for key, wid in (("pil_idx_fbmc", 0), ("pil_idx_ofdm", 1)):
            if any(s["wf"] == ("F", "O")[wid] for s in self.schemes.values()):
                keep[key] = np.ascontiguousarray(np.stack([d[key] for d in draws_list]), dtype=np.int32)
                st.pilot_idx[wid] = keep[key].ctypes.data
        st.on_device = 0
        return st, keep

    def run_batch(self, n_rep, n_iter, draws=None, seed=0, first_rep=0):
        """Returns err[rep, snr, it, scheme, csi, edge] (uint32)."""
        err = np.zeros((n_rep, self.n_snr, n_iter + 1, 3, 2, 2), dtype=np.uint32)
        dptr = C.byref(draws) if draws is not None else None
        self._check(self.lib.chest_run_batch(self._h, n_rep, n_iter, dptr, seed, first_rep, _ptr(err)))
        return err

    def run_batch_device(self, n_rep, n_iter, draws=None, seed=0, first_rep=0, err_dev_ptr=None):
        dptr = C.byref(draws) if draws is not None else None
        self._check(self.lib.chest_run_batch_device(self._h, n_rep, n_iter, dptr, seed, first_rep, err_dev_ptr))

    def run_batch_async(self, n_rep, n_iter, draws=None, seed=0, first_rep=0):
        """Enqueue the loop body and return; `wait()` collects the counts (one host thread, several contexts)."""
        dptr = C.byref(draws) if draws is not None else None
        self._check(self.lib.chest_run_batch_async(self._h, n_rep, n_iter, dptr, seed, first_rep))
        self._pending_shape = (n_rep, self.n_snr, n_iter + 1, 3, 2, 2)

    def wait(self):
        err = np.zeros(self._pending_shape, dtype=np.uint32)
        self._check(self.lib.chest_wait(self._h, _ptr(err)))
        return err

    def prefetch_draws(self, n_rep, host_draws):
        """Start the asynchronous upload of host draws; returns the device-side ChestDraws to pass to run_batch*."""
        dev = _lib.ChestDraws()
        self._check(self.lib.chest_prefetch_draws(self._h, n_rep, C.byref(host_draws), C.byref(dev)))
        return dev

    def generate_draws(self, n_rep, seed, first_rep=0):
        st = _lib.ChestDraws()
        self._check(self.lib.chest_generate_draws(self._h, n_rep, seed, first_rep, C.byref(st)))
        return st

    def download_draws(self, n_rep):
        """Host copies of the device-resident draws, as a list of per-realization dicts."""
        TP = self.T * self.paths
        du, pu = np.zeros((n_rep, TP)), np.zeros((n_rep, TP))
        noise = np.zeros((n_rep, self.n_snr, self.N), dtype=np.complex128)
        bits = {n: np.zeros((n_rep, s["n_bits"]), dtype=np.uint8) for n, s in self.schemes.items()}
        P = {("F", "O").index(s["wf"]): s["P"] for s in self.schemes.values()}
        pidx = {w: np.zeros((n_rep, p), dtype=np.int32) for w, p in P.items()}
        self._check(self.lib.chest_download_draws(
            self._h, n_rep, _ptr(du), _ptr(pu), _ptr(bits.get("aux")), _ptr(bits.get("cod")), _ptr(bits.get("ofdm")),
            _ptr(pidx.get(0)), _ptr(pidx.get(1)), _ptr(noise)))
        out = []
        for r in range(n_rep):
            d = dict(doppler_u=du[r].reshape(self.T, self.paths, order="F"),
                     phase_u=pu[r].reshape(self.T, self.paths, order="F"), noise=noise[r])
            for n in bits:
                d["bits_" + n] = bits[n][r]
            if 0 in pidx:
                d["pil_idx_fbmc"] = pidx[0][r]
            if 1 in pidx:
                d["pil_idx_ofdm"] = pidx[1][r]
            out.append(d)
        return out

    def bit_counts(self):
        nb = np.zeros((3, 2), dtype=np.int64)
        self._check(self.lib.chest_bit_counts(self._h, nb.ctypes.data_as(C.POINTER(C.c_int64))))
        return nb

    def draws_bytes(self, n_rep):
        return int(self.lib.chest_draws_bytes(self._h, n_rep))

    def get_state(self, what, name, rep, i_snr):
        s = self.schemes[name]
        n = {"y": s["K"], "hP": s["P"], "xD_est": s["n_data"], "xD_perf": s["n_data"], "hdiag": s["K"]}[what]
        code = {"y": 0, "hP": 1, "xD_est": 2, "xD_perf": 3, "hdiag": 4}[what]
        out = np.zeros(n, dtype=np.complex128)
        self._check(self.lib.chest_get_state(self._h, code, SCHEME_ID[name], rep, i_snr, _ptr(out)))
        return out

    # ------------------------------------------------------------------ measurement helpers
    def launch_count(self):
        return int(self.lib.chest_launch_count(self._h))

    def set_profiling(self, on=True):
        self._check(self.lib.chest_set_profiling(self._h, int(on)))

    def stage_times(self):
        ms = (C.c_float * 7)()
        self._check(self.lib.chest_stage_times(self._h, ms))
        return dict(zip(("draws", "k1_channel_tx", "k2_transmission_matrix", "k3_demod", "one_tap", "ic_iterations",
                         "total"), [float(x) for x in ms]))

    def banded_apply_stats(self):
        ms, by = C.c_float(0), C.c_double(0)
        self._check(self.lib.chest_banded_apply_stats(self._h, C.byref(ms), C.byref(by)))
        return float(ms.value), float(by.value)

    def event_record(self, slot):
        self._check(self.lib.chest_event_record(self._h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float(0)
        self._check(self.lib.chest_event_elapsed(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def n_units(self):
        """IC work units of the current batch size (3392 at the default configuration, B = 1024)."""
        n = C.c_int(0)
        self._check(self.lib.chest_unit_count(self._h, C.byref(n)))
        return n.value

    def set_perfect_csi_mode(self, mode):
        """'factored' (default): D never formed; 'dense': D = Q^H H G materialised per realization (K2)."""
        self._check(self.lib.chest_set_perfect_csi_mode(self._h, {"dense": 0, "factored": 1}[mode]))

    def set_mse_accumulation(self, on=True):
        """Also accumulate sum_i |h_est(i) - h(i)|^2 per (realization, SNR, iteration, scheme) in the next runs."""
        self._check(self.lib.chest_set_mse_accumulation(self._h, int(on)))

    def get_mse(self, n_rep, n_iter):
        """[rep, snr, it, scheme] sums of the last run_batch (scheme order aux, cod, ofdm)."""
        out = np.zeros((n_rep, self.n_snr, n_iter + 1, 3), dtype=np.float64)
        self._check(self.lib.chest_get_mse(self._h, out.ctypes.data))
        return out

    def set_precision(self, mode):
        """'fp64' (default): FP64 DMMA everywhere; 'split_bf16': the estimated-CSI cancellation on tcgen05 tensor cores with
        split-BF16 operands and FP32 accumulation in TMEM (the stated reduced-precision mode, ~1e-5 of the interference)."""
        self._check(self.lib.chest_set_precision(self._h, {"fp64": 0, "split_bf16": 1}[mode]))

    def precision_info(self):
        """(mode, dense BF16 flops one k_ic_est_tc launch executes, bytes of the packed W operand images)."""
        m, f, b = C.c_int(0), C.c_double(0), C.c_int64(0)
        self._check(self.lib.chest_precision_info(self._h, C.byref(m), C.byref(f), C.byref(b)))
        return ("fp64", "split_bf16")[m.value], f.value, b.value

    def set_estimator_mode(self, mode):
        """Form of the estimated-CSI cancellation (chest_b200.h): 'auto' (default: the factored form D_est = Q' H_est G only where
        it equals the thresholded W to rounding AND is cheaper), 'tiles' (always the thresholded W), 'factored' (every scheme
        with factors: the stated-tolerance mode, D_est within the removed 1e-8 entries of the reference's), 'factored_exact' (the
        factored form wherever it equals the thresholded W to rounding, i.e. CP-OFDM, whatever it costs)."""
        self._check(self.lib.chest_set_estimator_mode(self._h, {"auto": 0, "tiles": 1, "factored": 2, "factored_exact": 3}[mode]))

    def set_pseudo_channels(self, wf, M, removed_max=np.inf):
        """M[p, tap, n] = M_p(n, n - delay_tap): the pseudo-channels of the pilots (DS.m:260), taps = non-zero PDP entries."""
        M = np.ascontiguousarray(M, dtype=np.complex128)
        rm = 1e300 if not np.isfinite(removed_max) else float(removed_max)
        self._check(self.lib.chest_set_pseudo_channels(self._h, WF_ID[wf], M.shape[0], _ptr(M), rm))

    def set_estimator_factors(self, name, variant, R_inv, removed_max=np.inf):
        """pinv(R_hP_est) per SNR point (n_snr x P x P), as chest_build_mmse takes it."""
        Ri = np.ascontiguousarray(np.transpose(np.asarray(R_inv, dtype=np.complex128), (0, 2, 1)))     # column-major P x P per SNR point
        rm = 1e300 if not np.isfinite(removed_max) else float(removed_max)
        self._check(self.lib.chest_set_estimator_factors(self._h, SCHEME_ID[name], variant, Ri.shape[0], _ptr(Ri), rm))

    def estimator_info(self, name):
        """dict(mode, factored, removed_r, removed_w, ms): whether the scheme's estimated-CSI cancellation ran factored in the
        last batch, the largest magnitudes the 1e-8 thresholds removed at setup, device ms of the factored pass (profiling)."""
        m, f, rr, rw, ms = C.c_int(0), C.c_int(0), C.c_double(0), C.c_double(0), C.c_float(0)
        self._check(self.lib.chest_estimator_info(self._h, SCHEME_ID[name], C.byref(m), C.byref(f), C.byref(rr), C.byref(rw), C.byref(ms)))
        return dict(mode=("auto", "tiles", "factored", "factored_exact")[m.value], factored=bool(f.value), removed_r=rr.value, removed_w=rw.value,
                    ms=ms.value)

    def kernel_times(self):
        """Device ms of k_apply_hg, k_gemm_d, k_ic_main (sum), k_ic_light (sum) in the last profiled batch."""
        out = (C.c_float * 10)()
        self._check(self.lib.chest_kernel_times(self._h, out))
        return dict(k_apply_hg=out[0], k_gemm_d=out[1], k_ic_main=out[2], k_ic_light=out[3],
                    perfect_csi_chain=out[4], diag_d_gemm=out[5], k_synth_h=out[6], k_tx_symbols=out[7],
                    modulate_gemm=out[8], k_apply_h=out[9])

    def work_model(self, n_iter):
        out = (C.c_double * 8)()
        self._check(self.lib.chest_work_model(self._h, n_iter, out))
        return dict(k2_flops=out[0], est_flops=out[1], perf_flops=out[2], txdemod_flops=out[3],
                    w_bytes_per_ic_launch=out[4], precode_flops=out[5], est_main_flops=out[6],
                    factored_perf_flops=out[7])

    def fp64_peak(self, mode="dmma", iters=20000):
        t = C.c_double(0)
        self._check(self.lib.chest_fp64_peak(self._h, {"dmma": 0, "dfma": 1, "mix": 2}[mode], iters, C.byref(t)))
        return t.value


class MultiDevice:
    """chest_multi_*: one host thread driving one finalized, identically configured context per GPU."""

    def __init__(self, contexts):
        self.ctxs = list(contexts)
        self.lib = self.ctxs[0].lib
        hs = (C.c_uint64 * len(self.ctxs))(*[c._h.value for c in self.ctxs])
        m = C.c_uint64(0)
        self._m = None
        self.ctxs[0]._check(self.lib.chest_multi_create(hs, len(self.ctxs), C.byref(m)))
        self._m = m

    def run(self, n_rep_total, n_iter, seed=0, first_rep=0, want_err=True):
        """Returns (err[rep, snr, it, scheme, csi, edge] or None, totals[snr, it, scheme, csi, edge] uint64,
        device ms of the final NCCL all-reduce)."""
        c0 = self.ctxs[0]
        err = np.zeros((n_rep_total, c0.n_snr, n_iter + 1, 3, 2, 2), dtype=np.uint32) if want_err else None
        tot = np.zeros((c0.n_snr, n_iter + 1, 3, 2, 2), dtype=np.uint64)
        ms = C.c_float(0)
        c0._check(self.lib.chest_multi_run(self._m, n_rep_total, n_iter, seed, first_rep, _ptr(err), _ptr(tot), C.byref(ms)))
        return err, tot, float(ms.value)

    def close(self):
        if self._m is not None:
            self.lib.chest_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
