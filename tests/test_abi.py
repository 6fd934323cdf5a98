"""The C-ABI library loads without a GPU and exports every symbol include/gprb200.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

import gpr_jl_b200  # noqa: F401
from gpr_jl_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gprb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gprb_[a-z_A-Z0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = L.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not lib.has_symbol(n)]
    assert not missing, missing
    assert sorted(L.EXPORTS) == names  # the ctypes binding covers exactly the header


def test_version_and_error_string():
    lib = L.load_library()
    assert lib.dll.gprb_version() >= 100
    assert isinstance(lib.dll.gprb_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is checked on the CPU box")
    lib = L.load_library()
    h = C.c_void_p()
    rc = lib.dll.gprb_init(C.byref(h), 0)
    assert rc == -3  # GPRB_ERR_NODEVICE
    assert b"no CPU fallback" in lib.dll.gprb_last_error()
    with pytest.raises(L.GprbError):
        from gpr_jl_b200 import GP, SEArd, MeanZero
        import numpy as np
        GP(np.zeros((2, 4)), np.zeros(4), MeanZero(), SEArd([0.0, 0.0], 0.0))


def test_api_misuse_is_reported_not_crashed():
    lib = L.load_library()
    assert lib.dll.gprb_init(None, 0) == -1
    assert b"NULL" in lib.dll.gprb_last_error()
    assert lib.dll.gprb_lbfgs_selftest(0, 0, None, None, 1.0, None) == -1
