"""TEST INFRASTRUCTURE — makes tests/golden/ref_fusion_c1.npz: the point cloud the REFERENCE's own RunFusion
(DPE.cpp:1220-1370, run through oracle/_ref/ref_fusion_probe, which includes the reference sources where they lie)
produces from a fixed set of depth / normal / state maps, together with those maps.  RunFusion is CPU code, so this
runs in the build container (no GPU): `python oracle/make_fusion_golden.py`.

The maps are this implementation's own (CPU logic simulator, 5-view 160x120 c1 scene, whole schedule): what matters
for the fixture is only that both fusions read the same inputs.  tests/test_fusion_golden.py holds
oracle/fusion_oracle.cpp to the cloud (bit for bit, same order), and the device fusion is held to that oracle on the
GPU (tests/test_gpu_parity.py)."""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "dpe-mvs_b200", ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
import prep_cv2  # noqa: E402
import simpipe  # noqa: E402
import synth  # noqa: E402
from scenes import small_scene  # noqa: E402


def read_ply(path):
    raw = Path(path).read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    n = int([l for l in head.decode().split("\n") if l.startswith("element vertex")][0].split()[-1])
    rec = np.frombuffer(body, np.dtype([("xyz", "<f4", 3), ("bgr", "u1", 3)]))
    assert len(rec) == n
    return rec["xyz"].copy(), rec["bgr"].copy()


def main():
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "_ref/ref_fusion_probe"])
    spec, grays, cams, drs, pairs, gt = small_scene("c1", 0.25)
    V = len(grays)
    H, W = grays[0].shape
    state, _ = simpipe.run(grays, cams, drs, pairs, 2, seed=20261018)
    depth = np.stack([s["depth"] for s in state]).astype(np.float32)          # depths.dmb: out-of-range already zeroed
    normal = np.ascontiguousarray(np.stack([s["planes"][..., :3] for s in state]).astype(np.float32))   # (world normal, depth) planes
    st = np.stack([s["state"] for s in state]).astype(np.uint8)               # weak.bin: PixelState
    with tempfile.TemporaryDirectory() as tmp:
        folder = Path(tmp) / "scene"
        (folder / "images").mkdir(parents=True)
        (folder / "cams").mkdir()
        for v in range(V):
            with open(folder / "images" / f"{v:08d}.gray", "wb") as f:        # the shim's imread reads this sidecar
                f.write(np.array([H, W], np.int32).tobytes()); f.write(np.ascontiguousarray(grays[v]).tobytes())
            K, R, t = cams[v]
            synth.write_cam(folder / "cams" / f"{v:08d}_cam.txt", K, R, t, *drs[v])
            d = folder / "DPE" / f"{v:08d}"
            d.mkdir(parents=True)
            prep_cv2.write_dmb(d / "depths.dmb", depth[v])
            prep_cv2.write_dmb(d / "normals.dmb", normal[v])
            prep_cv2.write_dmb(d / "weak.bin", st[v])
        with open(folder / "pair.txt", "w") as f:
            f.write(f"{V}\n")
            for v in range(V):
                f.write(f"{v}\n{len(pairs[v])} " + " ".join(f"{j} 100.0" for j in pairs[v]) + "\n")
        subprocess.check_call([str(ROOT / "oracle" / "_ref" / "ref_fusion_probe"), str(folder)])
        xyz, bgr = read_ply(folder / "DPE" / "DPE.ply")
        # once more with <dense>/blocks/mask_<id>.jpg (DPE.cpp:1242-1268, 1296): grey masks, a reference pixel below 128 is skipped
        yy, xx = np.mgrid[0:H, 0:W]
        blocks = np.stack([np.where(((xx // 16 + yy // 12 + v) % 3) == 0, 40 + 20 * v, 130 + 25 * v).astype(np.uint8) for v in range(V)])
        blocks[:, :, :7] = 127; blocks[:, :, -7:] = 128            # both sides of the threshold
        (folder / "blocks").mkdir()
        for v in range(V):
            with open(folder / "blocks" / f"mask_{v}.gray", "wb") as f:       # the shim's imread reads the sidecar of mask_<id>.jpg
                f.write(np.array([H, W], np.int32).tobytes()); f.write(np.ascontiguousarray(blocks[v]).tobytes())
        subprocess.check_call([str(ROOT / "oracle" / "_ref" / "ref_fusion_probe"), str(folder)])
        xyz_b, bgr_b = read_ply(folder / "DPE" / "DPE.ply")
        # the cameras as the reference parsed them from the cam files (what both fusions must be fed)
        cam_rt = [synth.read_cam(folder / "cams" / f"{v:08d}_cam.txt") for v in range(V)]
    out = ROOT / "tests" / "golden" / "ref_fusion_c1.npz"
    np.savez_compressed(out, depth=depth, normal=normal, state=st,
                        gray=np.stack(grays), K=np.stack([c[0] for c in cam_rt]), R=np.stack([c[1] for c in cam_rt]),
                        t=np.stack([c[2] for c in cam_rt]), pairs=np.array(pairs, np.int32), ref_xyz=xyz, ref_bgr=bgr)
    print("wrote", out, "points", len(xyz), "bytes", out.stat().st_size)
    out_b = ROOT / "tests" / "golden" / "ref_fusion_c1_blocks.npz"      # inputs are those of ref_fusion_c1.npz
    np.savez_compressed(out_b, blocks=blocks, ref_xyz=xyz_b, ref_bgr=bgr_b)
    print("wrote", out_b, "points", len(xyz_b), "bytes", out_b.stat().st_size)


if __name__ == "__main__":
    main()
