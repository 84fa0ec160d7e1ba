"""Device time of the whole schedule over a 12-view c2-shape scene (1600x1200, 10 sources), views overlapped on the
context's streams as in production, per cost arithmetic.  usage: time_modes.py [modes...]  (2 exact, 1 fast, 0 centred)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT))
import capi
from bench import ensure_scene, load_scene_arrays, product_prep
modes = [int(a) for a in sys.argv[1:]] or [2, 1]
folder = ensure_scene("c2", 12, "c2v12")
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
sched = capi.stage_schedule(ns)
for rep in range(2):
    for m in modes:
        ctx.set_cost_arithmetic(m)
        g0 = ctx.stage_gpu_ms(); t0 = time.time()
        for (k, p) in sched:
            ctx.run_stage(k, p, 20261018); ctx.stage_commit()
        print(f"rep {rep} arithmetic {m}: gpu_ms {ctx.stage_gpu_ms() - g0:.1f} wall {time.time() - t0:.2f} s -> {V / ((ctx.stage_gpu_ms() - g0) * 1e-3):.3f} depth maps/s", flush=True)
