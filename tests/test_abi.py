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
