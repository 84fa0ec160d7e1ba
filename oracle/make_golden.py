"""TEST INFRASTRUCTURE — generates the golden vectors that pin oracle/ncc_oracle.py to the
reference's own device code.  Runs on a GPU box:

    python oracle/make_golden.py          # writes tests/golden/ref_probe_c1.npz

It renders a small synthetic scene, draws seeded pixels / plane hypotheses around the ground
truth, runs oracle/_ref/ref_probe (the reference's ComputeBilateralNCCOld and
ComputeGeomConsistencyCost, compiled from /root/reference and executed on the GPU) and stores
inputs + reference outputs.  tests/test_oracle_golden.py replays the inputs through the
float64 restatement on the CPU.
"""
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200"))
import synth  # noqa: E402


def camera_bytes(K, R, t, H, W, dmin, dmax):
    """struct Camera, main.h:50-59: K[9] R[9] t[3] c[3] height width depth_min depth_max."""
    K, R, t = (np.asarray(a, np.float32) for a in (K, R, t))
    c = -(R.astype(np.float64).T @ t.astype(np.float64)).astype(np.float32)   # DPE.cpp:362-367
    return K.tobytes() + R.tobytes() + t.tobytes() + c.tobytes() + struct.pack("<iiff", H, W, dmin, dmax)


def main():
    spec = synth.make_scene("c1", scale=0.25)     # 160 x 120
    W, H = spec.width, spec.height
    rv = [synth.render_view(spec, v, device="cpu") for v in range(4)]
    imgs = np.stack([r[0] for r in rv]).astype(np.float32)
    depths = np.stack([r[1] for r in rv]).astype(np.float32)
    depths[:, ::7, ::5] = 0.0                      # some invalid source depths (geom cost 3.0)
    cams = [tuple(np.asarray(a, np.float32) for a in spec.cams[v]) for v in range(4)]
    rng = np.random.default_rng(20261018)
    n = 600
    xy = np.stack([rng.integers(0, W, n), rng.integers(0, H, n)], 1).astype(np.int32)   # borders included
    K, R, t = cams[0]
    planes = np.zeros((n, 4), np.float32)
    for i, (x, y) in enumerate(xy):
        d = max(float(rv[0][1][y, x]), 0.5) * (1 + rng.normal(0, 0.02 if i % 3 else 0.3))
        nrm = R @ rv[0][2][y, x] + rng.normal(0, 0.1, 3)
        nrm /= np.linalg.norm(nrm)
        X = d * np.array([(x - K[0, 2]) / K[0, 0], (y - K[1, 2]) / K[1, 1], 1.0])
        planes[i] = [nrm[0], nrm[1], nrm[2], -float(nrm @ X)]
    blob = struct.pack("<iiii", W, H, 4, n) + imgs.tobytes() + depths.tobytes()
    for (Kc, Rc, tc) in cams:
        blob += camera_bytes(Kc, Rc, tc, H, W, 1.0, 10.0)
    blob += xy.tobytes() + planes.tobytes()
    tmp = Path("/tmp/ref_probe_in.bin")
    tmp.write_bytes(blob)
    subprocess.check_call([str(ROOT / "oracle" / "_ref" / "ref_probe"), str(tmp), "/tmp/ref_probe_out.bin"])
    out = np.fromfile("/tmp/ref_probe_out.bin", np.float32)
    ncc, geom = out[: n * 3].reshape(n, 3), out[n * 3:].reshape(n, 3)
    dst = ROOT / "tests" / "golden" / "ref_probe_c1.npz"
    np.savez_compressed(dst, images=imgs.astype(np.uint8), depths=depths,
                        K=np.stack([c[0] for c in cams]), R=np.stack([c[1] for c in cams]), t=np.stack([c[2] for c in cams]),
                        xy=xy, planes=planes, ref_ncc=ncc, ref_geom=geom)
    print("wrote", dst, "ncc mean", float(ncc.mean()), "frac 2.0", float((ncc == 2.0).mean()), "geom mean", float(geom.mean()))


if __name__ == "__main__":
    main()
