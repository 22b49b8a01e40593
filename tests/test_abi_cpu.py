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


def test_preprocess_engine_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    lib = capi.load()
    o = capi.SvinPreOptions()
    o.src_width, o.src_height, o.resize_factor, o.max_images = 64, 48, 1.0, 1
    ctx = C.c_void_p()
    assert lib.svin_pre_create(0, C.byref(o), C.byref(ctx)) == -3  # SVIN_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.svin_last_error()


def test_reference_arm_prints_one_json_line():
    # bench.py --impl reference: the CPU restatement on the host cores, same metric / unit / config keys as the GPU arm
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "keyframes/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def test_bench_gpu_arm_refuses_to_run_without_a_gpu():
    # the product arm of bench.py must not fall back to anything on a machine without CUDA
    import subprocess
    import sys
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pytest.skip("torch not importable")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
    assert p.stdout.strip() == ""          # and prints no JSON line
