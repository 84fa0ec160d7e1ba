"""The kernels' per-pixel logic under AddressSanitizer + UndefinedBehaviorSanitizer, on the CPU.

dpe_core.cuh / dpe_weak.cuh are __host__ __device__ code; oracle/dpe_hostsim.cu compiles them for the host (TEST
INFRASTRUCTURE, not a product path).  Built with the sanitizers (oracle/Makefile: _ref/libdpe_hostsim_asan.so) the
whole stage schedule runs with every array index, shift and conversion of the sweeps, the anchor search, the plane
fit, the weak update and the classifier checked — what compute-sanitizer would do on the GPU, without one.

usage: python tools/sim_sanitize.py            (re-executes itself with the sanitizer runtimes preloaded)
Scenes: the C1-shape scene cropped to a ragged 157 x 83 (tiles hanging over the image), and the weak-texture scene
(edge-mode sweeps, WEAK pixels, geometric-consistency stages); both cost arithmetics.  Prints one line per run and the
number of sanitizer reports (0 expected)."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SO = ROOT / "oracle" / "_ref" / "libdpe_hostsim_asan.so"

if os.environ.get("DPE_SIM_SANITIZE_CHILD") != "1":
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "_ref/libdpe_hostsim_asan.so"])
    pre = ":".join(subprocess.check_output(["gcc", f"-print-file-name={n}"], text=True).strip() for n in ("libasan.so", "libubsan.so"))
    env = dict(os.environ, LD_PRELOAD=pre, ASAN_OPTIONS="detect_leaks=0:halt_on_error=0", UBSAN_OPTIONS="print_stacktrace=1",
               DPE_SIM_SANITIZE_CHILD="1")
    p = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
    sys.stdout.write(p.stdout)
    reports = [l for l in p.stderr.splitlines() if "runtime error" in l or "ERROR: AddressSanitizer" in l]
    for l in reports[:40]:
        print(l)
    print(f"sanitizer reports: {len(reports)} (exit code {p.returncode})")
    sys.exit(1 if reports or p.returncode else 0)

for p in (ROOT / "tests", ROOT / "oracle", ROOT / "dpe-mvs_b200"):
    sys.path.insert(0, str(p))
import numpy as np
import capi
import hostsim
import simpipe
from scenes import small_scene

hostsim._lib = C.CDLL(str(SO))

spec, grays, cams, drs, pairs, gt = small_scene()
ragged = [np.ascontiguousarray(g[:83, :157]) for g in grays]
for exact in (0, 1):
    os.environ.pop("DPE_HOSTSIM_EXACT", None)
    if exact:
        os.environ["DPE_HOSTSIM_EXACT"] = "1"
    st, units = simpipe.run(ragged, cams, drs, pairs, 2, seed=7)
    print("c1 ragged 157x83, reference arithmetic" if exact else "c1 ragged 157x83, fast arithmetic", [int(u) for u in units], flush=True)

spec, grays, cams, drs, pairs, gt = small_scene("c4", 0.04, 4)
lib = capi.load()
H, W = grays[0].shape
sizes = simpipe.level_sizes(W, H, 2)
prep = []
for g in grays:
    per = []
    for k in range(2):
        e = np.empty((sizes[k][1], sizes[k][0]), np.uint8)
        l = np.empty((sizes[k][1], sizes[k][0]), np.int32)
        gg = np.ascontiguousarray(g)
        lib.dpe_host_problem_edges(gg.ctypes.data_as(C.c_void_p), W, H, 1 << (1 - k), e.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p))
        per.append((e, l))
    prep.append(per)
for exact in (0, 1):
    os.environ.pop("DPE_HOSTSIM_EXACT", None)
    if exact:
        os.environ["DPE_HOSTSIM_EXACT"] = "1"
    st, units = simpipe.run(grays, cams, drs, pairs, 2, prep=prep, seed=5)
    print(f"weak-texture {W}x{H}, " + ("reference" if exact else "fast") + " arithmetic", [int(u) for u in units],
          "WEAK pixels", sum(int((s["state"] == hostsim.WEAK).sum()) for s in st), flush=True)
