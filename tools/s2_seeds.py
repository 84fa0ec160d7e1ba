"""Which pixels differ from the reference kernels' dump after the first fine sweep (stage 6 replay, tests/golden/
ref_stage_weak.npz), and what did the reference read there?  For every differing pixel of the FIRST colour (the second
colour inherits differences from the first) the reference's new plane is looked up among the planes its direction-4
positions held before the sweep (s1) and after it (s2): a match with an after-sweep plane that is not a before-sweep
plane means the reference thread read a pixel another thread of the same launch had already written.  GPU box."""
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
race = int(sys.argv[1]) if len(sys.argv) > 1 else 2


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
    out = ctx.debug_read(7, (H, W, 4), np.float32).copy()
    acc = ctx.debug_read(11, (H, W), np.uint8).copy()
    st = ctx.debug_read(9, (H, W), np.uint8).copy()
    cost = ctx.debug_read(3, (H, W), np.float32).copy()
    ctx.close()
    return out, acc, st, cost


ours, acc, st, cost = run(race, 2)
s1, s2 = fx["s1_planes"], fx["s2_planes"]
b = lambda a: a.view(np.uint32)
diff = ~(b(ours) == b(s2)).all(-1)
yy, xx = np.mgrid[0:H, 0:W]
first = ((xx + yy) & 1) == 0          # BlackPixelUpdate runs first (DPE.cu:3199): black = (x + y) even
print(f"direction-4 mode {race}: {int(diff.sum())} of {diff.size} pixels differ after the first sweep; first colour {int((diff & first).sum())}, second {int((diff & ~first).sum())}")
for colour, name in ((first, "first"), (~first, "second")):
    kinds = {}
    for y, x in zip(*np.nonzero(diff & colour)):
        rp = b(s2[y, x]); op = b(ours[y, x])
        tag = "other"
        if (rp == b(s1[y, x])).all():
            tag = "reference kept its plane, ours changed"
        elif (op == b(s1[y, x])).all():
            tag = "ours kept its plane, reference changed"
        # where could the reference's plane have come from?
        src = []
        for kk in range(1, 80):
            qx, qy = x - kk, y - kk
            if qx < 0 or qy < 0:
                break
            if (rp == b(s2[qy, qx])).all() and not (rp == b(s1[qy, qx])).all():
                src.append(f"dir4 step {kk}: plane written in THIS launch")
            elif (rp == b(s1[qy, qx])).all():
                src.append(f"dir4 step {kk}: plane from before the launch")
        osrc = []
        for kk in range(1, 80):
            qx, qy = x - kk, y - kk
            if qx < 0 or qy < 0:
                break
            if (op == b(ours[qy, qx])).all() and not (op == b(s1[qy, qx])).all():
                osrc.append(f"dir4 step {kk} NEW")
            elif (op == b(s1[qy, qx])).all():
                osrc.append(f"dir4 step {kk} OLD")
        key = (tag, "ref:" + (src[0].split(":")[1].strip() if src else "not a direction-4 plane"))
        kinds[key] = kinds.get(key, 0) + 1
        if colour is first:
            print(f"  ({x},{y}) state {st[y, x]} edge {fx['edge'][y, x]} our code {acc[y, x]} | {tag} | ref source {src[:2]} | our source {osrc[:2]} | ref cost {fx['s2_costs'][y, x]:.7f} our cost {cost[y, x]:.7f} before {fx['s1_costs'][y, x]:.7f}")
    print(name, "colour:", kinds)

# ---- is every difference a race difference?  A first-colour pixel is "sensitive" when the direction-4 candidate it would
# pick from the maps as they were BEFORE the launch (s1) is not the one it would pick from the maps AFTER it (s2: other
# position, other plane, or none because the costs along the ray became NaN).  Iteration 0 on this scene: offset 5, step 2.
def pick(c, x, y):
    best, pos = np.float32(3.4e38), None
    for s in range(11):
        qx, qy = x - 5 - 2 * s, y - 5 - 2 * s
        if qx < 0 or qy < 0:
            break
        if best > c[qy, qx]:
            best, pos = c[qy, qx], (qx, qy)
    return pos


c1, c2 = fx["s1_costs"], fx["s2_costs"]
sens = np.zeros((H, W), bool)
for y in range(H):
    for x in range(y & 1, W, 2):
        if st[y, x] == 0:
            continue
        po, pn = pick(c1, x, y), pick(c2, x, y)
        if po is None and pn is None:
            continue
        sens[y, x] = (po is None) != (pn is None) or po != pn or not (b(s1[po[1], po[0]]) == b(s2[pn[1], pn[0]])).all()
d1 = diff & first
print(f"first-colour pixels whose direction-4 candidate was rewritten during the launch: {int(sens.sum())}; "
      f"differing first-colour pixels among them: {int((d1 & sens).sum())} of {int(d1.sum())} "
      f"(a differing pixel that is NOT among them would be an arithmetic difference: {int((d1 & ~sens).sum())})")
