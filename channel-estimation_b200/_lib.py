"""ctypes binding of include/chest_b200.h (the C-ABI shared library built from csrc/).

There is no fallback: if the library is missing, importing the symbols raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHEST_LIB", os.path.join(_HERE, "libchest_b200.so"))   # CHEST_LIB: development override

c_u64, c_i64, c_i32, c_int = C.c_uint64, C.c_int64, C.c_int32, C.c_int
p_d = C.POINTER(C.c_double)
p_u8 = C.POINTER(C.c_uint8)
p_i32 = C.POINTER(C.c_int32)
p_i64 = C.POINTER(C.c_int64)
p_u32 = C.POINTER(C.c_uint32)
p_f = C.POINTER(C.c_float)
vp = C.c_void_p


class ChestDraws(C.Structure):
    """struct chest_draws (chest_b200.h)."""
    _fields_ = [("doppler_u", vp), ("phase_u", vp), ("bits", vp * 3), ("pilot_idx", vp * 2),
                ("noise", vp), ("on_device", c_int), ("channel_gauss", vp)]


class ChestSvDraws(C.Structure):
    """struct chest_sv_draws (chest_b200.h)."""
    _fields_ = [("bits", vp * 3), ("pilot_idx", vp * 2), ("h", vp), ("noise", vp * 2), ("on_device", c_int)]


# name -> (restype, argtypes); every symbol include/chest_b200.h declares
SIGNATURES = {
    "chest_last_error": (C.c_char_p, []),
    "chest_device_info": (c_int, [c_int, C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int)]),
    "chest_create": (c_int, [c_int, C.POINTER(c_u64)]),
    "chest_destroy": (c_int, [c_u64]),
    "chest_set_channel": (c_int, [c_u64, c_int, c_int, vp, C.c_double, C.c_double, c_int, c_int]),
    "chest_set_waveform": (c_int, [c_u64, c_int, c_int, c_int, vp, vp]),
    "chest_set_constellation": (c_int, [c_u64, c_int, c_int, vp, vp]),
    "chest_set_scheme": (c_int, [c_u64, c_int, c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp,
                                 C.c_double, C.c_double, c_int, c_int, vp]),
    "chest_set_snr": (c_int, [c_u64, c_int, vp]),
    "chest_set_mmse": (c_int, [c_u64, c_int, c_int, c_int, vp, vp, vp]),
    "chest_setup_correlations": (c_int, [c_u64, c_int, c_int, vp, vp, C.c_double, vp, p_i64]),
    "chest_build_mmse": (c_int, [c_u64, c_int, c_int, c_int, vp, C.c_double]),
    "chest_release_setup": (c_int, [c_u64]),
    "chest_finalize": (c_int, [c_u64, c_int]),
    "chest_new_realization": (c_int, [c_u64, c_int, vp, vp]),
    "chest_new_realization_seeded": (c_int, [c_u64, c_int, c_u64, c_i64]),
    "chest_new_realization_gauss": (c_int, [c_u64, c_int, vp]),
    "chest_channel_info": (c_int, [c_u64, C.POINTER(c_int), p_d, C.POINTER(c_int)]),
    "chest_set_impulse_response": (c_int, [c_u64, c_int, vp]),
    "chest_get_impulse_response": (c_int, [c_u64, c_int, vp]),
    "chest_get_convolution_csc": (c_int, [c_u64, c_int, p_i64, vp, vp, vp]),
    "chest_convolve": (c_int, [c_u64, c_int, vp, c_int, vp]),
    "chest_transmission_matrix": (c_int, [c_u64, c_int, c_int, vp, vp]),
    "chest_modulate": (c_int, [c_u64, c_int, vp, c_int, vp]),
    "chest_demodulate": (c_int, [c_u64, c_int, vp, c_int, vp]),
    "chest_set_modem": (c_int, [c_u64, c_int, c_int, c_int, c_int, c_int, vp, c_int, c_int, c_int, c_int, vp, vp,
                                C.c_double, C.c_double]),
    "chest_transmission_matrix_batch": (c_int, [c_u64, c_int, c_int, p_f, p_d, vp]),
    "chest_transmission_matrix_entries": (c_int, [c_u64, c_int, c_int, c_int, vp, vp, vp]),
    "chest_modulate_fft": (c_int, [c_u64, c_int, vp, c_int, vp]),
    "chest_demodulate_fft": (c_int, [c_u64, c_int, vp, c_int, vp]),
    "chest_set_interpolation": (c_int, [c_u64, c_int, vp]),
    "chest_sv_run_batch": (c_int, [c_u64, c_int, vp, C.POINTER(ChestSvDraws), c_u64, c_i64, vp]),
    "chest_estimate": (c_int, [c_u64, c_int, c_int, c_int, vp, vp, vp]),
    "chest_draws_bytes": (c_i64, [c_u64, c_int]),
    "chest_run_batch": (c_int, [c_u64, c_int, c_int, C.POINTER(ChestDraws), c_u64, c_i64, vp]),
    "chest_run_batch_device": (c_int, [c_u64, c_int, c_int, C.POINTER(ChestDraws), c_u64, c_i64, vp]),
    "chest_run_batch_async": (c_int, [c_u64, c_int, c_int, C.POINTER(ChestDraws), c_u64, c_i64]),
    "chest_wait": (c_int, [c_u64, vp]),
    "chest_multi_create": (c_int, [C.POINTER(c_u64), c_int, C.POINTER(c_u64)]),
    "chest_multi_run": (c_int, [c_u64, c_i64, c_int, c_u64, c_i64, vp, vp, p_f]),
    "chest_multi_destroy": (c_int, [c_u64]),
    "chest_bit_counts": (c_int, [c_u64, p_i64]),
    "chest_generate_draws": (c_int, [c_u64, c_int, c_u64, c_i64, C.POINTER(ChestDraws)]),
    "chest_download_draws": (c_int, [c_u64, c_int, vp, vp, vp, vp, vp, vp, vp, vp]),
    "chest_get_state": (c_int, [c_u64, c_int, c_int, c_int, c_int, vp]),
    "chest_kernel_times": (c_int, [c_u64, vp]),
    "chest_set_perfect_csi_mode": (c_int, [c_u64, c_int]),
    "chest_set_mse_accumulation": (c_int, [c_u64, c_int]),
    "chest_get_mse": (c_int, [c_u64, vp]),
    "chest_set_precision": (c_int, [c_u64, c_int]),
    "chest_precision_info": (c_int, [c_u64, C.POINTER(c_int), p_d, p_i64]),
    "chest_set_estimator_mode": (c_int, [c_u64, c_int]),
    "chest_set_pseudo_channels": (c_int, [c_u64, c_int, c_int, vp, C.c_double]),
    "chest_set_estimator_factors": (c_int, [c_u64, c_int, c_int, c_int, vp, C.c_double]),
    "chest_estimator_info": (c_int, [c_u64, c_int, C.POINTER(c_int), C.POINTER(c_int), p_d, p_d, p_f]),
    "chest_unit_count": (c_int, [c_u64, C.POINTER(c_int)]),
    "chest_prefetch_draws": (c_int, [c_u64, c_int, C.POINTER(ChestDraws), C.POINTER(ChestDraws)]),
    "chest_launch_count": (c_i64, [c_u64]),
    "chest_set_profiling": (c_int, [c_u64, c_int]),
    "chest_stage_times": (c_int, [c_u64, p_f]),
    "chest_banded_apply_stats": (c_int, [c_u64, p_f, p_d]),
    "chest_work_model": (c_int, [c_u64, c_int, p_d]),
    "chest_event_record": (c_int, [c_u64, c_int]),
    "chest_event_elapsed": (c_int, [c_u64, c_int, c_int, p_f]),
    "chest_fp64_peak": (c_int, [c_u64, c_int, c_int, p_d]),
}

_lib = None


def load():
    """Load libchest_b200.so and attach the prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
