"""TEST INFRASTRUCTURE — adds the outputs of the reference's GenEdgeInform / FindNearestStrongPoint kernels
(DPE.cu:2483-2591, 2855-2889: edge_neigh, complex, label_boundary, weak_nearest_strong) for the stage-6 case of
tests/golden/ref_stage_weak.npz: the probe input is rebuilt from the inputs stored in that fixture, the reference's
own kernels run (oracle/_ref/ref_stage_probe, GPU box), and the four arrays are written to
tests/golden/ref_stage_weak_k2k3.npz (and gpurun_out/, which travels back).  The step-0 dump of the same run is
checked against the stored one, so the two fixtures describe the same run."""
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import capi  # noqa: E402
from make_stage_golden import run_probe, camera_bytes, FIELDS  # noqa: E402,F401


def main():
    fx = np.load(ROOT / "tests" / "golden" / "ref_stage_weak.npz")
    imgs = [i.astype(np.float32) for i in fx["images"]]
    n, H, W = fx["images"].shape
    cams = [(fx["K"][i], fx["R"][i], fx["t"][i]) for i in range(n)]
    k, p = capi.stage_schedule(2)[6]
    full_wh = tuple(int(x) for x in fx["full_wh"])
    dr = tuple(float(x) for x in fx["drange"])
    # the same blob run_probe writes, then the probe once more with the extras file
    out = run_probe("k2k3", imgs, cams, full_wh, dr, p, fx["prev_planes"], fx["prev_state"], fx["prev_selected"], list(fx["src_depths"]),
                    fx["edge"], fx["edge_low"], fx["label"])
    assert np.array_equal(out[0]["state"], fx["s0_state"]) and np.array_equal(out[0]["neighbours"], fx["s0_neighbours"]), "not the run of the stored fixture"
    ex = Path("/tmp/stage_k2k3_extras.bin")
    subprocess.check_call([str(ROOT / "oracle" / "_ref" / "ref_stage_probe"), "/tmp/stage_k2k3_in.bin", "/tmp/stage_k2k3_out2.bin", str(ex)])
    raw = ex.read_bytes()
    P = W * H
    off = 0
    en = np.frombuffer(raw, np.int16, P * 16, off).reshape(H, W, 8, 2).copy(); off += P * 32
    cx = np.frombuffer(raw, np.float32, P, off).reshape(H, W).copy(); off += P * 4
    lb = np.frombuffer(raw, np.int16, P * 16, off).reshape(H, W, 8, 2).copy(); off += P * 32
    ns = np.frombuffer(raw, np.int16, P * 2, off).reshape(H, W, 2).copy(); off += P * 4
    assert off == len(raw)
    for dst in (ROOT / "tests" / "golden", ROOT / "gpurun_out"):
        dst.mkdir(exist_ok=True)
        np.savez_compressed(dst / "ref_stage_weak_k2k3.npz", s0_edge_neigh=en, s0_complex=cx, s0_label_boundary=lb, s0_nearest_strong=ns)
    print("wrote ref_stage_weak_k2k3.npz", en.shape, cx.shape, lb.shape, ns.shape)


if __name__ == "__main__":
    main()
