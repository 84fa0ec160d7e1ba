"""After the first strong sweep of one fine stage of a full-size scene: which pixels differ from the reference's own kernels
(oracle/_ref/ref_stage_probe, same inputs), and is each of them explained by the direction-4 race?  The direction-4
candidate of a pixel (both passes of the edge-adaptive sampling, DPE.cu:1250-1343) is re-derived in numpy from the
costs BEFORE the launch and from the costs AFTER it; a first-colour pixel whose candidate is the same plane either way cannot
have been touched by the race — if it differs from the reference, the difference is arithmetic or logic.  GPU box.
usage: sweep_seeds_scene.py <config> <scale> <views> <stage 4..7> [race mode]"""
import ctypes as C
import os
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
config, scale, n_views, target = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
race = int(sys.argv[5]) if len(sys.argv) > 5 else 2
import capi, hostsim, simpipe  # noqa: E402
import make_stage_golden as msg  # noqa: E402
from scenes import small_scene  # noqa: E402
spec, grays, cams, drs, pairs, gt = small_scene(config, scale, n_views)
lib = capi.load()
H, W = grays[0].shape
sizes = simpipe.level_sizes(W, H, 2)
prep = []
for g in grays:
    per = []
    for k in range(2):
        e = np.empty((sizes[k][1], sizes[k][0]), np.uint8)
        l = np.empty((sizes[k][1], sizes[k][0]), np.int32)
        gg = np.ascontiguousarray(g)
        lib.dpe_host_problem_edges(gg.ctypes.data_as(C.c_void_p), W, H, 1 << (1 - k), e.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p))
        per.append((e, l))
    prep.append(per)
sched = capi.stage_schedule(2)
v = len(grays) - 1            # the last view: in profile mode its scratch arrays stay readable after the stage
ids = [v] + list(pairs[v])
ctx = capi.Context(0)
capi.upload_scene(ctx, grays, cams, drs, pairs, 2)
for vv in range(len(grays)):
    for kk in range(2):
        ctx.set_prep(vv, kk, *prep[vv][kk])
ctx.set_reference_race(race)
ctx.set_profile(len(grays))
for si in range(target):
    ctx.run_stage(*sched[si], msg.SEED); ctx.stage_commit()
k, p = sched[target]
prev_k = sched[target - 1][0]
maps = [ctx.get_maps(i, prev_k) for i in range(len(grays))]
pm = maps[v]
w, h = sizes[k]
planes = np.concatenate([pm["normal"], pm["depth"][..., None]], -1).astype(np.float32)
state, sel = pm["state"], pm["selected"]
assert prev_k == k, "stage 4 changes scale; use stages 5..7"
src_d = [maps[i]["depth"] for i in pairs[v]] if p.geom_consistency else None
imgs = [grays[i].astype(np.float32) for i in ids]
edge, edge_low, label = prep[v][k][0], prep[v][0][0], prep[v][k][1]
dumps = msg.run_probe(f"seeds{target}", imgs, [cams[i] for i in ids], (W, H), drs[v], p, planes, state, sel, src_d, edge, edge_low, label)
ctx.debug_stop_after(2)
ctx.run_stage(k, p, msg.SEED)
ours = ctx.debug_read(7, (h, w, 4), np.float32).copy()
acc = ctx.debug_read(11, (h, w), np.uint8).copy()
st = ctx.debug_read(9, (h, w), np.uint8).copy()
en = ctx.debug_read(12, (h, w, 8, 2), np.int16).copy()
ctx.close()
s1, s2 = dumps[1]["planes"], dumps[2]["planes"]
c1, c2 = dumps[1]["costs"], dumps[2]["costs"]
b = lambda a: a.view(np.uint32)
diff = ~(b(ours) == b(s2)).all(-1)
yy, xx = np.mgrid[0:h, 0:w]
first = ((xx + yy) & 1) == 0
print(f"{config} {w}x{h} stage {target}, direction-4 mode {race}: {int(diff.sum())} of {diff.size} pixels differ after the first sweep "
      f"({100 * diff.mean():.3f} %); first colour {int((diff & first).sum())}, second {int((diff & ~first).sum())}")
max_edge = np.float32(max(h, w)) / np.float32(30.0)


def picks(c, x, y):
    """direction 4 = (-1,-1), iteration 0 (offset 5): pass 1 (edge-adaptive) and pass 2 (11 steps of 2) positions"""
    out = []
    on_edge = edge[y, x] != 0
    ex, ey = int(en[y, x, 4, 0]), int(en[y, x, 4, 1])
    dist = np.float32(np.sqrt(float((ex - x) ** 2 + (ey - y) ** 2)) / 1.4142135623730951)
    if on_edge:
        dist = np.float32(22.0)
    elif ex == -1 or ey == -1 or dist > max_edge:
        dist = np.float32(max_edge / 1.4142135623730951)
    step_num = min(max(11, int(dist / np.float32(2))), 22)
    step_len = max(int(dist / np.float32(step_num)), 2)
    for (n, sl) in ((step_num, step_len),) + (() if on_edge else ((11, 2),)):
        best, pos = np.float32(3.4e38), None
        for s in range(n):
            qx, qy = x - 5 - s * sl, y - 5 - s * sl
            if qx < 0 or qy < 0:
                break
            if best > c[qy, qx]:
                best, pos = c[qy, qx], (qx, qy)
        out.append(pos)
    return out


d1 = diff & first
ys, xs = np.nonzero(d1)
clean = []
for y, x in zip(ys, xs):
    po, pn = picks(c1, x, y), picks(c2, x, y)
    sens = False
    for a_, b_ in zip(po, pn):
        if (a_ is None) != (b_ is None):
            sens = True
        elif a_ is not None and (a_ != b_ or not (b(s1[a_[1], a_[0]]) == b(s2[b_[1], b_[0]])).all()):
            sens = True
    if not sens:
        clean.append((x, y))
print(f"differing first-colour pixels: {len(ys)}; with a direction-4 candidate rewritten during the launch: {len(ys) - len(clean)}; "
      f"NOT explained by the race: {len(clean)}")
for (x, y) in clean[:40]:
    print(f"   ({x},{y}) state {st[y, x]} on-edge {edge[y, x] != 0} our code {acc[y, x]} edge_neigh4 {en[y, x, 4].tolist()} ours {ours[y, x].tolist()} ref {s2[y, x].tolist()} "
          f"ref cost {c2[y, x]:.7f} before {c1[y, x]:.7f} ref changed {not (b(s1[y, x]) == b(s2[y, x])).all()} ours changed {not (b(s1[y, x]) == b(ours[y, x])).all()}")
# where could the planes of the unexplained pixels have come from?  every candidate position of the 8 directions (both passes)
DIRS = [(0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (1, 1), (-1, 1), (1, -1)]


def all_positions(x, y):
    out = []
    on_edge = edge[y, x] != 0
    for d, (dx, dy) in enumerate(DIRS):
        ex, ey = int(en[y, x, d, 0]), int(en[y, x, d, 1])
        dist = np.float32(np.sqrt(float((ex - x) ** 2 + (ey - y) ** 2)))
        if d >= 4:
            dist = np.float32(dist / 1.4142135623730951)
        if on_edge:
            dist = np.float32(22.0)
        elif ex == -1 or ey == -1 or dist > max_edge:
            dist = np.float32(max_edge)
            if d >= 4:
                dist = np.float32(dist / 1.4142135623730951)
        step_num = min(max(11, int(dist / np.float32(2))), 22)
        step_len = max(int(dist / np.float32(step_num)), 2)
        if d < 4 and step_len % 2 == 1:
            step_len -= 1
        fx = fy = 0
        if d > 4:
            if d % 2:
                fx = dx
            else:
                fy = dy
        for ps, (n, sl) in enumerate(((step_num, step_len),) + (() if on_edge else ((11, 2),))):
            for s_ in range(n):
                qx, qy = x + 5 * dx + s_ * sl * dx + fx, y + 5 * dy + s_ * sl * dy + fy
                if 0 <= qx < w and 0 <= qy < h:
                    out.append((d, ps, s_, qx, qy))
    return out


for (x, y) in clean[:10]:
    pos = all_positions(x, y)
    hit_r = [(d, ps, s_, qx, qy, float(c1[qy, qx])) for (d, ps, s_, qx, qy) in pos if (b(s1[qy, qx]) == b(s2[y, x])).all() or (b(s2[qy, qx]) == b(s2[y, x])).all()]
    hit_o = [(d, ps, s_, qx, qy, float(c1[qy, qx])) for (d, ps, s_, qx, qy) in pos if (b(s1[qy, qx]) == b(ours[y, x])).all()]
    best = {}
    for (d, ps, s_, qx, qy) in pos:
        k_ = (d, ps)
        if k_ not in best or c1[qy, qx] < best[k_][0]:
            best[k_] = (float(c1[qy, qx]), qx, qy, s_)
    print(f"   ({x},{y}): reference's new plane found at candidate positions (dir, pass, step, x, y, cost before): {hit_r[:4]}; ours at {hit_o[:4]}")
    print(f"        min-cost pick per (dir, pass) from the costs before the launch: { {k_: v for k_, v in sorted(best.items())} }")
    nanc = [(d, ps, s_) for (d, ps, s_, qx, qy) in pos if not np.isfinite(c1[qy, qx])]
    print(f"        non-finite costs among the candidates before the launch: {nanc[:6]}; selected-views of the pixel before / ref after: {int(dumps[1]['selected'][y, x]):#x} / {int(dumps[2]['selected'][y, x]):#x}")

# how many first-colour pixels were exposed to the race at all?
exposed = 0
sample = [(x, y) for y in range(0, h, 3) for x in range(y & 1, w, 6) if st[y, x] != 0]
for (x, y) in sample:
    po, pn = picks(c1, x, y), picks(c2, x, y)
    for a_, b_ in zip(po, pn):
        if (a_ is None) != (b_ is None) or (a_ is not None and (a_ != b_ or not (b(s1[a_[1], a_[0]]) == b(s2[b_[1], b_[0]])).all())):
            exposed += 1
            break
print(f"exposure: {exposed} of {len(sample)} sampled first-colour pixels have a direction-4 candidate that the launch rewrote")
np.savez_compressed(ROOT / "gpurun_out" / f"sweep_seeds_{config}_{w}x{h}_stage{target}.npz", ours=ours, acc=acc, st=st, en=en, s1=s1, s2=s2, c1=c1, c2=c2, edge=edge,
                    sel1=dumps[1]["selected"], sel2=dumps[2]["selected"])
