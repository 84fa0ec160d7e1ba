"""A/B of the warp-cooperative sweep + classifier (csrc/dpe_coop.cuh) against the per-pixel forms
(DPE_VARIANT_PER_PIXEL_COSTS) on one resident scene: device time of the whole schedule, the per-class kernel
times, and whether the two forms leave bit-identical maps at full size.

usage: python tools/ab_coop.py [config=c2] [n_views=12] [scale=1.0] [timed passes per form=2]
The scene comes from $DPE_BENCH_DIR (bench.py's cache) and is rendered on the CPU when it is not there."""
import json
import os
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT))
import capi
from bench import ensure_scene, load_scene_arrays, product_prep

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 12
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
n_pass = int(sys.argv[4]) if len(sys.argv) > 4 else 2
PER_PIXEL = 8
SEED = 20261018

t0 = time.time()
folder = ensure_scene(cfg, n_views, f"{cfg}v{n_views}s{scale}", scale)
grays, cams, drs, pairs = load_scene_arrays(folder)
V = len(grays); H, W = grays[0].shape
lib = capi.load()
ctx = capi.Context(0)
ns = capi.compute_round_num(W, H)
ctx.scene_begin(V, W, H, ns)
for v in range(V):
    ctx.set_view(v, grays[v], *cams[v], *drs[v]); ctx.set_pairs(v, pairs[v])
    for k, (e, l) in enumerate(product_prep(lib, grays[v], ns)):
        ctx.set_prep(v, k, e, l)
ctx.commit()
ctx.set_cost_arithmetic({"fast": 1, "centred": 0}.get(os.environ.get("DPE_ARITH", ""), 2))
ctx.set_reference_race(2)
sched = capi.stage_schedule(ns)
print(f"scene {cfg} views {V} size {W}x{H} sources {len(pairs[0])} stages {len(sched)} set-up {time.time() - t0:.1f} s", flush=True)


def one_pass():
    m0 = ctx.stage_gpu_ms()
    for (k, p) in sched:
        ctx.run_stage(k, p, SEED); ctx.stage_commit()
    return ctx.stage_gpu_ms() - m0


def digest():
    out = []
    for v in range(V):
        m = ctx.get_maps(v, ns - 1)
        out.append({k: np.ascontiguousarray(m[k]).copy() for k in ("depth", "normal", "state", "selected")})
    return out


ctx.debug_set_variants(PER_PIXEL)
one_pass()                                  # warm: pools, every view has a depth map
res = {"config": cfg, "views": V, "width": W, "height": H, "sources": len(pairs[0]), "forms": {}}
maps = {}
for name, variants in (("per_pixel", PER_PIXEL), ("cooperative", 0)):
    ctx.debug_set_variants(variants)
    ms = [one_pass() for _ in range(n_pass)]
    maps[name] = digest()
    ctx.set_profile(V)
    one_pass()
    prof = ctx.get_profile()
    ctx.set_profile(0)
    res["forms"][name] = {"variants": variants, "schedule_ms": ms, "maps_per_s": V / (min(ms) * 1e-3),
                          "kernel_ms": {k: round(c["ms"], 3) for k, c in prof.items()},
                          "kernel_Gunits_per_s": {k: round(c["units"] / c["ms"] / 1e6, 2) for k, c in prof.items() if c["units"]}}
    print(name, json.dumps(res["forms"][name]), flush=True)
same = {}
for key in ("depth", "normal", "state", "selected"):
    eq = []
    for a, b in zip(maps["per_pixel"], maps["cooperative"]):
        x, y = a[key], b[key]
        if x.dtype == np.float32:
            x, y = x.view(np.uint32), y.view(np.uint32)
        eq.append(float((x == y).mean()))
    same[key] = min(eq)
res["bit_identical_fraction_min_over_views"] = same
a, b = res["forms"]["per_pixel"], res["forms"]["cooperative"]
res["speedup_schedule"] = min(a["schedule_ms"]) / min(b["schedule_ms"])
res["speedup_by_class"] = {k: round(a["kernel_ms"][k] / b["kernel_ms"][k], 3) for k in ("strong_sweep", "classify_refine") if k in b["kernel_ms"]}
print(json.dumps(res))
