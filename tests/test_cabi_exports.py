"""CPU: the C-ABI library builds, loads, and exports every symbol include/nvbio_b200.h declares."""
import ctypes as C
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nvbio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nvb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_everything():
    from nvbio_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    lib.nvb_version.restype = C.c_int
    assert lib.nvb_version() == 100
    lib.nvb_error_string.restype = C.c_char_p
    assert b"invalid" in lib.nvb_error_string(C.c_int(-1))


def test_python_mirror_lists_same_exports():
    from nvbio_b200 import _lib
    assert set(_lib.EXPORTS) == set(declared_symbols())


def test_argument_validation_without_gpu():
    """invalid arguments are rejected before any CUDA call"""
    from nvbio_b200 import _lib
    L = _lib.lib()
    assert L.nvb_fm_match(None, None, C.c_uint32(1), C.c_uint32(0), None, None) == -1
    tb = C.c_size_t(0)
    assert L.nvb_banded_gotoh_score(C.c_int(4), C.c_int(1), None, None, None, None, C.c_uint32(1), None, None, None, C.byref(tb), None) == -1


def test_no_oracle_in_product():
    """the product package must not import / link the oracle"""
    pkg = os.path.join(ROOT, "nvbio_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "synth.py", (f, "mentions the oracle")
