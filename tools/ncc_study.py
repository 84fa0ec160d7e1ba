"""Tap-loop study on the B200: candidate-scoring throughput (G NCC units/s) of the variants in
dpe_kernels.cu:launch_ncc_bench on converged maps of the 12-view c2-shape scene."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT))
import capi
from bench import ensure_scene, load_scene_arrays, product_prep
folder = ensure_scene("c2", 12, "c2v12")
grays, cams, drs, pairs = load_scene_arrays(folder)
V = len(grays); H, W = grays[0].shape
lib = capi.load(); ctx = capi.Context(0)
ns = capi.compute_round_num(W, H)
ctx.scene_begin(V, W, H, ns)
for v in range(V):
    ctx.set_view(v, grays[v], *cams[v], *drs[v]); ctx.set_pairs(v, pairs[v])
    for k, (e, l) in enumerate(product_prep(lib, grays[v], ns)):
        ctx.set_prep(v, k, e, l)
ctx.commit()
for (k, p) in capi.stage_schedule(ns):
    ctx.run_stage(k, p, 20261018); ctx.stage_commit()
names = {0: "as-is/4", 1: "rows6/3"}
out = []
for arith, aname in ((1, "fast"), (2, "reference, operation by operation")):
    ctx.set_cost_arithmetic(arith)
    for var in sorted(names):
        best, cs = 0.0, None
        for view in (0, 5):
            r, c = ctx.bench_ncc(view, var, 8, 3)
            best = max(best, r); cs = c if view == 0 else cs
        out.append(dict(arithmetic=aname, variant=names[var], gunits=best / 1e9, checksum=cs))
        print(out[-1], flush=True)
(ROOT / "gpurun_out" / "ncc_study.json").write_text(json.dumps(out, indent=1))
