"""The C-ABI library loads, exports every symbol include/dpe_b200.h declares, and fails loudly
without a GPU (no CPU fallback).  CPU only, no compute calls."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

import capi

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "dpe_b200.h").read_text()
    return sorted(set(re.findall(r"DPE_API\s+[\w\s\*]+?\b(dpe_\w+)\s*\(", text)))


def test_header_symbols_exported():
    lib = capi.load()
    syms = header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dpe_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == syms, "capi.SYMBOLS must list exactly the header's entry points"


def test_stage_params_layout_matches_header():
    assert C.sizeof(capi.StageParams) == 9 * 4


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.DpeError):
        capi.Context(0)
    lib = capi.load()
    assert lib.dpe_run_pipeline(b"/nonexistent", 0, 0, 0, 0, 1, 0, 0, 0) != 0


def test_python_surface_signature():
    import inspect
    import DPE_MVS
    sig = inspect.signature(DPE_MVS.dpe_mvs)
    assert list(sig.parameters) == ["dense_folder", "gpu_index", "verbose", "fusion", "viz", "depth", "normal", "weak", "edge"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["gpu_index"] == 0 and d["verbose"] is True and d["fusion"] is False and d["viz"] is False
    assert d["depth"] is True and d["normal"] is False and d["weak"] is False and d["edge"] is False
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="DPE-MVS failed with code"):
            DPE_MVS.dpe_mvs("/nonexistent_dense_folder", verbose=False)


def test_cli_usage():
    exe = ROOT / "dpe-mvs_b200" / "bin" / "DPE"
    assert exe.exists()
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode != 0 and "USAGE: DPE dense_folder" in p.stderr


def test_schedule_matches_reference():
    # main.cpp:508-567 for R = 3
    s = capi.stage_schedule(3)
    assert len(s) == 12
    assert [k for k, _ in s] == [0] * 4 + [1] * 4 + [2] * 4
    assert [p.state for _, p in s] == [0, 2, 2, 2, 1, 2, 2, 2, 1, 2, 2, 2]
    assert [p.geom_consistency for _, p in s] == [0, 1, 1, 1] * 3
    assert [p.use_apd for _, p in s] == [0] * 4 + [1] * 8
    assert [p.weak_peak_radius for _, p in s] == [6, 4, 2, 2] * 3
    assert [p.rotate_time for _, p in s[4:]] == [2] * 4 + [4] * 4
    assert abs(s[4][1].ransac_threshold - 0.00875) < 1e-7 and abs(s[8][1].ransac_threshold - 0.0075) < 1e-7
    assert capi.compute_round_num(640, 480) == 2 and capi.compute_round_num(1600, 1200) == 2
    assert capi.compute_round_num(3024, 2016) == 3 and capi.compute_round_num(1920, 1080) == 3


def test_every_entry_point_declares_its_arguments_and_survives_null():
    """ctypes passes an undeclared Python int as a 32-bit C int: a pointer handed to an entry point without argtypes is
    cut in half.  Every symbol of the header must have its argtypes declared in capi.py, and — called with a null context
    and null / zero arguments, in a child process — must come back with an error code instead of crashing."""
    lib = capi.load()
    assert [n for n in capi.SYMBOLS if getattr(lib, n).argtypes is None] == []
    child = r"""
import sys, ctypes as C
sys.path.insert(0, %r)
import capi
lib = capi.load()
ints = (C.c_int, C.c_long, C.c_size_t, C.c_uint64, C.c_uint32, C.c_longlong, C.c_ulonglong)
for name in capi.SYMBOLS:
    f = getattr(lib, name)
    args = [t(0) if t in ints or t in (C.c_float, C.c_double) else None for t in f.argtypes]
    print(name, flush=True)
    f(*args)
print("ALL-RETURNED")
""" % str(ROOT / "dpe-mvs_b200")
    p = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "ALL-RETURNED" in p.stdout, (p.returncode, p.stdout.splitlines()[-1:], p.stderr[-300:])
