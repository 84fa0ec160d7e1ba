"""Test helper: the reference's stage schedule (main.cpp:508-567) driven over the TEST-ONLY CPU
logic simulator (oracle/hostsim.py), Jacobi order like the product (all views of a stage
read the previous stage's depth maps)."""
import sys
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "dpe-mvs_b200", ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))
import hostsim  # noqa: E402


def level_sizes(W, H, n_scales):
    out = []
    for k in range(n_scales):
        f = np.float32(1.0) / np.float32(1 << (n_scales - 1 - k))
        out.append((int(np.floor(float(np.float32(W) * f) + 0.5)), int(np.floor(float(np.float32(H) * f) + 0.5))))
    return out


def run(grays, cams, depth_ranges, pairs, n_scales, prep=None, seed=20261018, stages=None, quant=1, views=None):
    """grays: list of HxW uint8; prep[v][k] = (edge, label) for scale index k (0 = coarsest).
    Returns per-view dicts of the last stage + list of per-stage eval units."""
    V = len(grays)
    H, W = grays[0].shape
    sizes = level_sizes(W, H, n_scales)
    pyr = [[hostsim.resize_linear(g.astype(np.float32), *sizes[k]) if sizes[k] != (W, H) else g.astype(np.float32)
            for k in range(n_scales)] for g in grays]
    state = [None] * V
    depths = {}
    units = []
    sched = hostsim.stage_schedule(n_scales)
    if stages is not None:
        sched = sched[:stages]
    todo = list(range(V)) if views is None else views
    for si, (k, p) in enumerate(sched):
        new_state, new_depths, u = [None] * V, {}, 0.0
        for v in todo:
            ids = [v] + list(pairs[v])
            imgs = [pyr[i][k] for i in ids]
            cc = [cams[i] for i in ids]
            sd = [depths[i] for i in pairs[v]] if p.geom_consistency else None
            prev = None if state[v] is None else (state[v]["planes"], state[v]["state"], state[v]["selected"])
            e = l = el = None
            if prep is not None:
                e, l = prep[v][k]
                el = prep[v][0][0]
            r = hostsim.run_stage(imgs, cc, depth_ranges[v], (W, H), p, seed, view=v, stage_counter=si, prev=prev,
                                  src_depths=sd, edge=e, edge_low=el, label=l, quant=quant)
            new_state[v] = r
            new_depths[v] = r["depth"]
            u += r["units"]
        state, depths = new_state, new_depths
        units.append(u)
    return state, units
