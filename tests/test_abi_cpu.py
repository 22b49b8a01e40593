"""CPU checks on the product library: it loads, exports every symbol include/svin_b200.h declares,
and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from svin_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "svin_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(svin_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = capi.load()
    syms = _header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/svin_b200.h but not exported"
    assert set(capi.EXPORTED_SYMBOLS) <= set(syms)
    assert b"svin_b200" in lib.svin_version()


def test_default_options_match_ceres_defaults():
    lib = capi.load()
    o = capi.SvinBaOptions()
    lib.svin_ba_default_options(C.byref(o))
    assert o.max_num_iterations == 10 and o.initial_trust_region_radius == 1e4
    assert o.function_tolerance == 1e-6 and o.min_relative_decrease == 1e-3 and o.jacobi_scaling == 1


def test_engine_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    lib = capi.load()
    ctx = C.c_void_p()
    rc = lib.svin_ba_create(0, C.byref(ctx))
    assert rc == -3  # SVIN_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.svin_last_error()
