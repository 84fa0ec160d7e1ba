"""First GPU contact: texture probes, kernel-vs-simulator cost check, strong-path stages on
c1 and on a 12-view slice of the c2 shape, with timings.  Writes gpurun_out/first.json."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200"))
import capi, synth, hostsim

out = {}
OUT = ROOT / "gpurun_out"; OUT.mkdir(exist_ok=True)
ctx = capi.Context(0)
w = ctx.probe_tex_weights(4096)
np.save(OUT / "tex_weights.npy", w)
q = np.round(w * 256)
out["tex_weights_all_multiples_of_1_256"] = bool(np.allclose(w * 256, q, atol=1e-6))
fr = np.arange(4097) / 4096.0
out["tex_weights_round_match"] = float((np.floor(fr * 256 + 0.5) / 256 == w).mean())
out["tex_weights_trunc_match"] = float((np.floor(fr * 256) / 256 == w).mean())
out["tex_rate_taps_per_s"] = [ctx.probe_tex_rate(2048, 2048, 200) for _ in range(3)]
out["tex_rate_small_tex"] = ctx.probe_tex_rate(256, 256, 200)
out["fma_rate_per_s"] = [ctx.probe_fma_rate(4000) for _ in range(3)]
print(json.dumps(out), flush=True)

def load(cfgname, scale=1.0, n_views=None):
    spec = synth.make_scene(cfgname, scale=scale, n_views=n_views)
    t0 = time.time()
    rv = [synth.render_view(spec, v) for v in range(spec.n_views)]
    print("render", cfgname, time.time() - t0, flush=True)
    imgs = [r[0] for r in rv]
    cams = [tuple(np.asarray(a, np.float32) for a in c) for c in spec.cams]
    drs = []
    for r in rv:
        valid = r[1][r[1] > 0]
        drs.append((float(np.percentile(valid, 1)) * 0.75, float(np.percentile(valid, 99)) * 1.25))
    return spec, rv, imgs, cams, drs, synth.select_pairs(spec)

def score(maps, gt_d, gt_n):
    d = maps["depth"]; m = gt_d > 0
    rel = np.abs(d - gt_d) / np.maximum(gt_d, 1e-6)
    ang = np.degrees(np.arccos(np.clip((maps["normal"] * gt_n).sum(-1), -1, 1)))
    st = maps["state"]
    return dict(within1=float((rel[m] < 0.01).mean()), within5=float((rel[m] < 0.05).mean()),
                n5deg=float((ang[m] < 5).mean()), zero=float((d == 0).mean()),
                state=[float((st == k).mean()) for k in range(3)])

# ---- c1: cost parity kernel vs simulator + full strong-only schedule
spec, rv, imgs, cams, drs, pairs = load("c1")
ns = capi.upload_scene(ctx, imgs, cams, drs, pairs)
W, H = spec.width, spec.height
rng = np.random.default_rng(0)
npx = 4000
xy = np.stack([rng.integers(8, W - 8, npx), rng.integers(8, H - 8, npx)], 1).astype(np.int32)
gt_d = rv[0][1]; gt_n = rv[0][2]
K, R, t = cams[0]
planes = np.zeros((npx, 4), np.float32)
for i, (x, y) in enumerate(xy):
    d = gt_d[y, x] * (1 + rng.normal(0, 0.01))
    n = R @ gt_n[y, x]
    n = n + rng.normal(0, 0.05, 3); n /= np.linalg.norm(n)
    X = d * np.array([(x - K[0, 2]) / K[0, 0], (y - K[1, 2]) / K[1, 1], 1.0])
    planes[i] = [n[0], n[1], n[2], -float(n @ X)]
g0 = ctx.cost_eval(0, ns - 1, xy, planes, len(pairs[0]), mode=0)
g1 = ctx.cost_eval(0, ns - 1, xy, planes, len(pairs[0]), mode=1)
ids = [0] + pairs[0]
fimgs = [imgs[i].astype(np.float32) for i in ids]; fcams = [cams[i] for i in ids]
for qm in (0, 1, 2):
    s = hostsim.cost_eval(fimgs, fcams, (W, H), xy, planes, quant=qm)
    out[f"cost_gpu_hw_vs_sim_q{qm}"] = [float(np.abs(g0 - s).max()), float(np.median(np.abs(g0 - s)))]
    out[f"cost_gpu_exact_vs_sim_q{qm}"] = [float(np.abs(g1 - s).max()), float(np.median(np.abs(g1 - s)))]
np.savez(OUT / "cost_c1.npz", xy=xy, planes=planes, g0=g0, g1=g1)
print(json.dumps(out), flush=True)

def run_schedule(ctx, ns, n_views, tag, strong_only=True, seed=7):
    ctx.set_count_evals(True)
    res = []
    for (k, p) in capi.stage_schedule(ns):
        if strong_only:
            p.use_apd = 0
        u0, m0, l0 = ctx.eval_units(), ctx.stage_gpu_ms(), ctx.kernel_launches()
        t0 = time.time()
        ctx.run_stage(k, p, seed)
        ctx.stage_commit()
        dt = time.time() - t0
        res.append(dict(scale=k, state=p.state, geom=p.geom_consistency, wall_s=dt, gpu_ms=ctx.stage_gpu_ms() - m0,
                        units=ctx.eval_units() - u0, launches=ctx.kernel_launches() - l0))
        print(tag, res[-1], flush=True)
    return res

out["c1_stages"] = run_schedule(ctx, ns, spec.n_views, "c1")
out["c1_view0"] = score(ctx.get_maps(0, ns - 1), gt_d, gt_n)
print(json.dumps(out["c1_view0"]), flush=True)
# without eval counting (timing)
ctx2 = capi.Context(0)
ns = capi.upload_scene(ctx2, imgs, cams, drs, pairs)
t0 = time.time()
for (k, p) in capi.stage_schedule(ns):
    p.use_apd = 0
    ctx2.run_stage(k, p, 7); ctx2.stage_commit()
out["c1_total_wall_s_nocount"] = time.time() - t0
out["c1_total_gpu_ms_nocount"] = ctx2.stage_gpu_ms()
ctx2.close()

# ---- c2 shape, 12 views, 10 sources each
spec, rv, imgs, cams, drs, pairs = load("c2", n_views=12)
ns = capi.upload_scene(ctx, imgs, cams, drs, pairs)
out["c2_12v_stages"] = run_schedule(ctx, ns, spec.n_views, "c2")
out["c2_12v_view0"] = score(ctx.get_maps(0, ns - 1), rv[0][1], rv[0][2])
out["c2_12v_view5"] = score(ctx.get_maps(5, ns - 1), rv[5][1], rv[5][2])
np.save(OUT / "c2_depth0.npy", ctx.get_maps(0, ns - 1)["depth"])
ctx3 = capi.Context(0)
ns = capi.upload_scene(ctx3, imgs, cams, drs, pairs)
t0 = time.time()
for (k, p) in capi.stage_schedule(ns):
    p.use_apd = 0
    ctx3.run_stage(k, p, 7); ctx3.stage_commit()
out["c2_12v_total_wall_s_nocount"] = time.time() - t0
out["c2_12v_total_gpu_ms_nocount"] = ctx3.stage_gpu_ms()
out["c2_12v_depth_maps_per_s_strong_only"] = 12 / (time.time() - t0)
(OUT / "first.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
