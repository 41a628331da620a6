"""The C-ABI library loads and exports every symbol include/chest_b200.h declares."""
import ctypes
import os
import re

import chest_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "chest_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chest_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(chest_b200._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in chest_b200._lib.SIGNATURES, "no ctypes prototype for " + n
    assert sorted(chest_b200._lib.SIGNATURES) == names


def test_error_reporting_without_context():
    lib = chest_b200._lib.load()
    assert lib.chest_destroy(0) == 0
    assert lib.chest_launch_count(0) == 0
    assert lib.chest_set_snr(0, 0, None) == -1
    assert b"argument check failed" in lib.chest_last_error()


def test_mex_shim_compiles_against_stub():
    import subprocess
    shim = os.path.join(ROOT, "matlab", "chest_mex.c")
    subprocess.run(["gcc", "-c", "-Wall", "-Werror", "-I", os.path.join(ROOT, "matlab", "stub"),
                    "-I", os.path.join(ROOT, "include"), "-o", "/tmp/chest_mex_test.o", shim], check=True)


def test_matlab_wrappers_use_only_gateway_commands_that_exist():
    """Static cross-check of the MATLAB side (it cannot be executed here): every chest_mex('command', ...) the classdef
    wrappers issue is dispatched by matlab/chest_mex.c, and every wrapper class north_star names is present."""
    import glob
    import re
    shim = open(os.path.join(ROOT, "matlab", "chest_mex.c")).read()
    known = set(re.findall(r'strcmp\(cmd, "([a-z_]+)"\)', shim))
    used = set()
    files = glob.glob(os.path.join(ROOT, "matlab", "+*", "*.m"))
    for f in files:
        used |= set(re.findall(r"chest_mex\('([a-z_]+)'", open(f).read()))
    assert used and used <= known, sorted(used - known)
    names = {os.path.relpath(f, os.path.join(ROOT, "matlab")) for f in files}
    for want in ("+Channel/FastFading.m", "+Modulation/FBMC.m", "+Modulation/OFDM.m", "+Modulation/SignalConstellation.m",
                 "+ChannelEstimation/PilotSymbolAidedChannelEstimation.m",
                 "+ChannelEstimation/ImaginaryInterferenceCancellationAtPilotPosition.m", "+ChestB200/Simulation.m"):
        assert want in names, want
    # every ABI entry point the MATLAB tier needs is reachable through the gateway
    for sym in ("chest_set_modem", "chest_modulate_fft", "chest_demodulate_fft", "chest_setup_correlations", "chest_build_mmse",
                "chest_multi_run", "chest_run_batch_async", "chest_wait", "chest_set_impulse_response", "chest_sv_run_batch",
                "mexAtExit"):
        assert sym in shim, sym
