"""Where does the first sweep of stage 0 diverge from the reference kernels?  (GPU box; tests/golden/ref_stage_first.npz)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT / "tests"))
import capi
from scenes import small_scene
fx = np.load(ROOT / "tests" / "golden" / "ref_stage_first.npz")
spec, grays, cams, drs, pairs, gt = small_scene("c4", 0.05, 5)
ch, cw = fx["images"].shape[1:]
ctx = capi.Context(0)
capi.upload_scene(ctx, grays, cams, drs, pairs, 2, active=(0, 1))
ctx.set_cost_arithmetic(2); ctx.set_view_order(1)
k, p = capi.stage_schedule(2)[0]
def run(step):
    ctx.debug_stop_after(step); ctx.run_stage(k, p, 20261018); ctx.stage_commit()
    return (ctx.debug_read(7, (ch, cw, 4), np.float32), ctx.debug_read(3, (ch, cw), np.float32), ctx.debug_read(8, (ch, cw), np.uint32))
pl, co, se = run(1)
img = ctx.debug_read(10, (ch, cw), np.float32)
print("coarse image of view 0 bitwise equal to the fixture's:", (img == fx["images"][0]).mean(), "max |d|", np.abs(img - fx["images"][0]).max())
d1 = np.abs(co - fx["s1_costs"]); d1 = d1[np.isfinite(d1) & (d1 > 0)]
print("step1 nonzero |dcost|:", len(d1), np.percentile(d1, [10, 50, 90]) if len(d1) else "")
print("step1: planes bitwise", (pl == fx["s1_planes"]).all(-1).mean(), "costs bitwise", (co == fx["s1_costs"]).mean(),
      "max |dcost|", np.nanmax(np.abs(co - fx["s1_costs"])), "selected", (se == fx["s1_selected"]).mean())
bad = np.argwhere(co != fx["s1_costs"])
for (y, x) in bad[:10]:
    print("  s1 cost differs", x, y, co[y, x], fx["s1_costs"][y, x], hex(se[y, x]), hex(fx["s1_selected"][y, x]))
pl, co, se = run(2)
acc = ctx.debug_read(11, (ch, cw), np.uint8)
rp, rc, rs = fx["s2_planes"], fx["s2_costs"], fx["s2_selected"]
same = (pl == rp).all(-1)
print("step2: planes bitwise", same.mean(), "costs bitwise", ((co == rc) | (np.isnan(co) & np.isnan(rc))).mean(), "selected", (se == rs).mean())
print("  among pixels with the same plane: costs bitwise", ((co == rc) | (np.isnan(co) & np.isnan(rc)))[same].mean(), "selected equal", (se == rs)[same].mean())
cd = same & ~((co == rc) | (np.isnan(co) & np.isnan(rc)))
print("  same plane, different cost:", int(cd.sum()), "pixels; |dcost| / cost percentiles 10/50/90:", np.percentile(np.abs(co - rc)[cd] / np.abs(rc)[cd], [10, 50, 90]))
print("  accepted-candidate histogram, all pixels:", np.bincount(acc.ravel(), minlength=15).tolist())
print("  accepted-candidate histogram, same plane / different cost:", np.bincount(acc[cd], minlength=15).tolist())
for (y, x) in np.argwhere(cd)[:12]:
    print(f"    ({x},{y}) colour {(x+y)%2} ours {co[y,x]:.9g} ref {rc[y,x]:.9g} rel {abs(co[y,x]-rc[y,x])/abs(rc[y,x]):.3g} sel {se[y,x]:x}/{rs[y,x]:x} plane changed in sweep: {not (pl[y,x] == fx['s1_planes'][y,x]).all()}")
# which integer view weights (15 draws) reproduce each stored cost from the per-view costs of the stored plane?
import itertools
pts = np.argwhere(cd)[:8]
if len(pts):
    xy = np.array([[x, y] for (y, x) in pts], np.int32)
    pv = ctx.cost_eval(0, 0, xy, np.stack([pl[y, x] for (y, x) in pts]), len(pairs[0]))
    for (y, x), c in zip(pts, pv):
        best = {}
        for w in itertools.product(range(16), repeat=len(c)):
            if sum(w) != 15: continue
            acc = np.float32(0)
            for wi, ci in zip(w, c):
                if wi > 0: acc = np.float32(np.float32(wi) * ci + acc)
            val = np.float32(acc / np.float32(15))
            for name, tgt in (("ours", co[y, x]), ("ref", rc[y, x])):
                d = abs(float(val) - float(tgt))
                if name not in best or d < best[name][0]: best[name] = (d, w)
        print(f"    ({x},{y}) per-view costs {c.tolist()} -> ours weights {best['ours'][1]} (|d| {best['ours'][0]:.2g}), ref weights {best['ref'][1]} (|d| {best['ref'][0]:.2g})")
bad = np.argwhere(~same)
dn = np.abs(pl[..., :3] - rp[..., :3]).max(-1)
tiny = (~same) & (dn < 1e-4) & (np.abs(pl[..., 3] - rp[..., 3]) < 1e-4 * np.abs(rp[..., 3]))
print("  differing pixels:", len(bad), "of which near-identical (1e-4):", int(tiny.sum()))
colour = (bad[:, 0] + bad[:, 1]) % 2
print("  by checkerboard colour:", np.bincount(colour, minlength=2).tolist())
dc = (co - rc)[~same]
print("  cost ours - ref at differing pixels: <-1e-3:", int((dc < -1e-3).sum()), " |d|<=1e-3:", int((np.abs(dc) <= 1e-3).sum()), " >1e-3:", int((dc > 1e-3).sum()), " nan:", int(np.isnan(dc).sum()))
# is the reference's plane a copy of one of OUR neighbours' planes (propagation) or new (refinement)?
s1 = fx["s1_planes"]
def source(plane, y, x, field):
    for yy in range(max(0, y - 30), min(ch, y + 31)):
        for xx in range(max(0, x - 30), min(cw, x + 31)):
            if (field[yy, xx] == plane).all():
                return (xx - x, yy - y)
    return None
for (y, x) in bad[:40]:
    src_ref = source(rp[y, x], y, x, s1) or source(rp[y, x], y, x, rp)
    src_our = source(pl[y, x], y, x, s1) or source(pl[y, x], y, x, pl)
    print(f"  ({x},{y}) colour {(x+y)%2} ours n=({pl[y,x,0]:.4f},{pl[y,x,1]:.4f},{pl[y,x,2]:.4f}) d={pl[y,x,3]:.5f} c={co[y,x]:.6f} sel={se[y,x]:x} from {src_our} | ref n=({rp[y,x,0]:.4f},{rp[y,x,1]:.4f},{rp[y,x,2]:.4f}) d={rp[y,x,3]:.5f} c={rc[y,x]:.6f} sel={rs[y,x]:x} from {src_ref}")
