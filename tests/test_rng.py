"""The XORWOW restatement (dpe-mvs_b200/csrc/dpe_rng.h) against cuRAND itself: oracle/curand_check.cu
compiles the CUDA toolkit's device-API header for the host and compares curand_init(seed, y, x) states,
curand() and curand_uniform() draws with ours, bit for bit.  CPU only."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_xorwow_matches_curand(tmp_path):
    exe = tmp_path / "curand_check"
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-w", str(ROOT / "oracle" / "curand_check.cu"), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr
