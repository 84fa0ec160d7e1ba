"""How much of the fine-stage difference to the reference is the direction-4 race?  Replays the stored
stage-6 inputs (tests/golden/ref_stage_weak.npz) several times with the reference's racy sampling positions and with
the race-free default and compares the runs with each other and with the reference kernels' dump, after the first
strong sweep (step 2) and after the last (step 10).  GPU box."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200"))
import capi
fx = np.load(ROOT / "tests" / "golden" / "ref_stage_weak.npz")
imgs = fx["images"]; n, H, W = imgs.shape
dr = tuple(float(x) for x in fx["drange"])
k, p = capi.stage_schedule(2)[6]


def run(race, step):
    ctx = capi.Context(0)
    ctx.scene_begin(n, W, H, 2)
    for i in range(n):
        ctx.set_view(i, imgs[i], fx["K"][i], fx["R"][i], fx["t"][i], *dr)
        ctx.set_pairs(i, list(range(1, n)) if i == 0 else [])
    ctx.set_prep(0, 1, fx["edge"], fx["label"])
    ctx.set_prep(0, 0, fx["edge_low"], np.full(fx["edge_low"].shape, -1, np.int32))
    ctx.set_active(0, 1)
    ctx.commit()
    ctx.set_view_order(1); ctx.set_reference_race(race)
    ctx.debug_set_maps(0, 1, fx["prev_planes"], fx["prev_state"], fx["prev_selected"])
    for j in range(1, n):
        ctx.debug_set_maps(j, 1, atlas_depth=fx["src_depths"][j - 1])
    ctx.debug_stop_after(step)
    ctx.run_stage(k, p, 20261018)
    out = ctx.debug_read(7, (H, W, 4), np.float32)
    if step == 2:
        global last_accept, last_state
        last_accept = ctx.debug_read(11, (H, W), np.uint8); last_state = ctx.debug_read(9, (H, W), np.uint8)
    ctx.close()
    return out


for step in (2, 10):
    for race in (1, 0):
        runs = [run(race, step) for _ in range(4)]
        same = [float((runs[0] == r).all(-1).mean()) for r in runs[1:]]
        ref = float((runs[0] == fx[f"s{step}_planes"]).all(-1).mean())
        print(f"step {step} ref_race={race}: run 0 vs runs 1-3 bit-identical planes {same}; run 0 vs reference dump {ref:.4f}", flush=True)

pl = run(1, 2)
diff = ~(pl == fx["s2_planes"]).all(-1)
print("step 2, ref_race=1: pixels differing from the reference dump:", int(diff.sum()), "of", diff.size)
print("  accepted-candidate histogram, all pixels:      ", np.bincount(last_accept.ravel(), minlength=15).tolist())
print("  accepted-candidate histogram, differing pixels:", np.bincount(last_accept[diff], minlength=15).tolist())
print("  state of differing pixels (0 weak, 1 strong, 2 unknown):", np.bincount(last_state[diff], minlength=3).tolist(), " on an edge:", int((fx["edge"][diff] > 0).sum()))
chg_ref = ~(fx["s2_planes"] == fx["s1_planes"]).all(-1); chg_our = ~(pl == fx["s1_planes"]).all(-1)
print("  differing pixels: reference changed its plane in the sweep", int((chg_ref & diff).sum()), ", ours did", int((chg_our & diff).sum()), ", both", int((chg_ref & chg_our & diff).sum()))
ys, xs = np.nonzero(diff)
for y, x in list(zip(ys, xs))[:25]:
    print(f"    ({x},{y}) code {last_accept[y, x]} edge {fx['edge'][y, x]} ours {pl[y, x].tolist()} ref {fx['s2_planes'][y, x].tolist()} ref cost {fx['s2_costs'][y, x]:.7f} s1 cost {fx['s1_costs'][y, x]:.7f}")
