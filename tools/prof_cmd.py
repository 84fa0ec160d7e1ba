"""Short, deterministic workload for ncu: 2 reference views of a 12-view c2-shape scene
(1600x1200, 10 sources), all 8 stages, single context.  Used for the launch list and the
--set full capture of the dominant kernel (B200_PROFILING.md)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT))
import capi
from bench import ensure_scene, load_scene_arrays, product_prep

n_prof = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2"
n_views = int(sys.argv[3]) if len(sys.argv) > 3 else 12
variants = int(sys.argv[4]) if len(sys.argv) > 4 else 0      # dpe_debug_set_variants
scale = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
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
ctx.debug_set_variants(variants)
sched = capi.stage_schedule(ns)
for (k, p) in sched:                      # warm pass: every view gets real depth maps
    ctx.run_stage(k, p, 20261018); ctx.stage_commit()
# ncu --profile-from-start off: dpe_run_stage brackets the profiled views with
# cudaProfilerStart/Stop, so only they are captured; the other views of each stage run after them
ctx.set_profile(n_prof)
import os
if os.environ.get("DPE_ARITH"):
    ctx.set_cost_arithmetic(int(os.environ["DPE_ARITH"]))
t0 = time.time()
for (k, p) in sched:
    ctx.run_stage(k, p, 20261018)
    ctx.stage_commit()
print("config", cfg, "views", V, "size", W, H, "variants", variants)
print("wall", time.time() - t0, "gpu_ms", ctx.stage_gpu_ms(), "launches", ctx.kernel_launches())
for k, v in ctx.get_profile().items():
    print(f"{k:16s} ms={v['ms']:10.3f} launches={v['launches']:4d} units={v['units']:.4g}" + (f"  Gunits/s={v['units']/v['ms']/1e6:.2f}" if v['units'] else ""))
