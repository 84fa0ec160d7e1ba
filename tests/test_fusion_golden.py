"""The fusion checker (oracle/fusion_oracle.cpp, a CPU restatement of RunFusion, DPE.cpp:1220-1370) against the point
cloud the REFERENCE's own RunFusion made from the same maps (tests/golden/ref_fusion_c1.npz, produced by
oracle/make_fusion_golden.py through oracle/_ref/ref_fusion_probe, which includes the reference's sources where they
lie): same number of points, same order, coordinates and colours bit for bit.  This pins the oracle the device fusion
is held to on the GPU (tests/test_gpu_parity.py::test_device_fusion_matches_cpu_oracle).  CPU only."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def oracle_fuse(depth, normal, state, bgr, K, R, t, pairs, blocks=None):
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "_ref/libfusion_oracle.so"])
    lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libfusion_oracle.so"))
    lib.fusion_oracle_run.restype = C.c_long
    lib.fusion_oracle_run_blocks.restype = C.c_long
    V, H, W = depth.shape
    src = np.ascontiguousarray(pairs, np.int32)
    cap = V * H * W
    xyz = np.empty((cap, 3), np.float32); out_bgr = np.empty((cap, 3), np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    maps = [np.ascontiguousarray(a) for a in (depth, normal, state, bgr)]
    cam = [np.ascontiguousarray(a) for a in (K.reshape(V, 9), R.reshape(V, 9), t.reshape(V, 3))]
    if blocks is None:
        n = lib.fusion_oracle_run(V, W, H, *[vp(a) for a in maps + cam], vp(src), src.shape[1], vp(xyz), vp(out_bgr), C.c_long(cap))
    else:
        bl = np.ascontiguousarray(blocks, np.uint8)
        n = lib.fusion_oracle_run_blocks(V, W, H, *[vp(a) for a in maps], vp(bl), *[vp(a) for a in cam], vp(src), src.shape[1], vp(xyz), vp(out_bgr), C.c_long(cap))
    return xyz[:n], out_bgr[:n]


def test_fusion_oracle_reproduces_the_reference_cloud():
    fx = np.load(ROOT / "tests" / "golden" / "ref_fusion_c1.npz")
    gray = fx["gray"]
    bgr = np.repeat(gray[..., None], 3, axis=-1)          # the reference build read grey sidecars as colour images
    xyz, col = oracle_fuse(fx["depth"], fx["normal"], fx["state"], bgr, fx["K"].astype(np.float32), fx["R"].astype(np.float32),
                           fx["t"].astype(np.float32), fx["pairs"])
    ref_xyz, ref_bgr = fx["ref_xyz"], fx["ref_bgr"]
    assert len(ref_xyz) > 5000
    assert len(xyz) == len(ref_xyz), (len(xyz), len(ref_xyz))
    assert np.array_equal(xyz.view(np.uint32), ref_xyz.view(np.uint32)), float((xyz != ref_xyz).any(-1).mean())
    assert np.array_equal(col, ref_bgr)


def test_fusion_oracle_reproduces_the_reference_cloud_with_block_masks():
    """<dense>/blocks/mask_<id>.jpg (DPE.cpp:1242-1268, 1296-1298): a reference pixel whose mask value is below 128 is not
    fused as a reference pixel (it may still serve as a source pixel of another view)."""
    fx = np.load(ROOT / "tests" / "golden" / "ref_fusion_c1.npz")
    fb = np.load(ROOT / "tests" / "golden" / "ref_fusion_c1_blocks.npz")
    bgr = np.repeat(fx["gray"][..., None], 3, axis=-1)
    xyz, col = oracle_fuse(fx["depth"], fx["normal"], fx["state"], bgr, fx["K"].astype(np.float32), fx["R"].astype(np.float32),
                           fx["t"].astype(np.float32), fx["pairs"], blocks=fb["blocks"])
    ref_xyz, ref_bgr = fb["ref_xyz"], fb["ref_bgr"]
    assert 5000 < len(ref_xyz) < len(fx["ref_xyz"])
    assert len(xyz) == len(ref_xyz), (len(xyz), len(ref_xyz))
    assert np.array_equal(xyz.view(np.uint32), ref_xyz.view(np.uint32))
    assert np.array_equal(col, ref_bgr)
