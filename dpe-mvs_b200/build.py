"""In-tree build of the native code (nvcc / g++ only; no cmake, no JIT cache).

Artifacts (all git-ignored, all travel to the GPU box with the snapshot):
  dpe-mvs_b200/lib/libdpe_b200.so       CUDA kernels (sm_100a) + C ABI + C++ host pipeline
  dpe-mvs_b200/DPE_MVS/_dpe.*.so         pybind11 module (drop-in for the reference's _dpe)
  dpe-mvs_b200/bin/DPE                   CLI (drop-in for the reference's ./DPE)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "lib"
BIN = HERE / "bin"
OBJ = HERE / "build"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# experiments: DPE_BUILD_TAG=<tag> DPE_BUILD_DEFS="-DDPE_CTAS_PER_SM=5" builds lib/libdpe_b200_<tag>.so beside the
# product library (capi.py loads it when DPE_LIB points at it)
TAG = os.environ.get("DPE_BUILD_TAG", "")
DEFS = os.environ.get("DPE_BUILD_DEFS", "").split()
if TAG:
    OBJ = HERE / f"build_{TAG}"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(str(c) for c in cmd), flush=True)
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(str(c) for c in cmd[:6])} ...")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def _headers():
    return list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("host/*.h")) + [
        HERE.parent / "include" / "dpe_b200.h"]


def build_lib(verbose=False, force=False) -> Path:
    LIB.mkdir(exist_ok=True)
    OBJ.mkdir(exist_ok=True)
    hdrs = _headers()
    objs = []
    cu = [CSRC / "dpe_kernels.cu", CSRC / "dpe_capi.cu", CSRC / "dpe_fusion.cu"]
    cpp = sorted((CSRC / "host").glob("*.cpp"))
    for src in cu + cpp:
        o = OBJ / (src.stem + ".o")
        if force or _newer(o, [src] + hdrs):
            extra = ["-Xptxas", "-v"] if (verbose and src.suffix == ".cu") else []
            if src.name == "dpe_kernels.cu":
                # the reference is built with --use_fast_math (csrc/DPE-MVS/CMakeLists.txt:16): approximate
                # division / sqrt / exp / sin / cos and flush-to-zero.  The kernels follow the reference's
                # expressions, so the same flag gives the same kind of rounding in the same places.
                extra = extra + ["--use_fast_math"]
            if src.name == "jpeg_luma.cpp":
                # the IDCT's 32-bit intermediates may wrap on damaged files; wrap-around must be defined behaviour there
                extra = extra + ["-Xcompiler", "-fwrapv"]
            _run([NVCC, *ARCH, *COMMON, *extra, *DEFS, "-I", CSRC, "-I", HERE.parent / "include", "-c", src, "-o", o], verbose)
        objs.append(o)
    so = LIB / (f"libdpe_b200_{TAG}.so" if TAG else "libdpe_b200.so")
    if force or _newer(so, objs):
        _run([NVCC, *ARCH, "-shared", "-o", so, *objs, "-lnvjpeg", "-lcudart", "-lpthread", "-ldl"], verbose)
    return so


def build_pybind(verbose=False, force=False) -> Path:
    import pybind11
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    out = HERE / "DPE_MVS" / f"_dpe{suffix}"
    src = CSRC / "bindings.cpp"
    lib = build_lib(verbose, force)
    if force or _newer(out, [src, lib]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I", pybind11.get_include(),
              "-I", sysconfig.get_paths()["include"], "-I", HERE.parent / "include", src, "-o", out,
              f"-L{LIB}", "-ldpe_b200", "-Wl,-rpath,$ORIGIN/../lib"], verbose)
    return out


def build_cli(verbose=False, force=False) -> Path:
    BIN.mkdir(exist_ok=True)
    out = BIN / "DPE"
    src = CSRC / "main.cpp"
    lib = build_lib(verbose, force)
    if force or _newer(out, [src, lib]):
        _run(["g++", "-O2", "-std=c++17", "-I", HERE.parent / "include", src, "-o", out, f"-L{LIB}", "-ldpe_b200",
              "-Wl,-rpath,$ORIGIN/../lib"], verbose)
    return out


def build_all(verbose=False, force=False):
    build_lib(verbose, force)
    if (CSRC / "bindings.cpp").exists():
        build_pybind(verbose, force)
    if (CSRC / "main.cpp").exists():
        build_cli(verbose, force)


if __name__ == "__main__":
    build_all(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print("build ok")
