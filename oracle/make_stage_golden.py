"""TEST INFRASTRUCTURE — golden per-step states of ONE (view, stage) computed by the reference's own
kernels (oracle/_ref/ref_stage_probe, built from /root/reference), for two cases:

  first   stage 0 (FIRST_INIT, coarse scale, no weak path, no geometric consistency)
  weak    stage 6 (REFINE_ITER at the fine scale: weak/edge path + geometric consistency), started from
          the maps this implementation produced for stages 0..5 on the GPU

Runs on a GPU box:   python oracle/make_stage_golden.py      -> tests/golden/ref_stage_{first,weak}.npz
tests/test_stage_golden.py replays the stored inputs through the CPU logic simulator step by step.
"""
import ctypes as C
import os
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
import capi, hostsim, synth  # noqa: E402
from scenes import small_scene  # noqa: E402
import simpipe  # noqa: E402

SEED = 20261018          # = DPE_REF_SEED of oracle/cvshim (the probe's curand seed)
FIELDS = [("planes", np.float32, 4), ("costs", np.float32, 1), ("selected", np.uint32, 1), ("state", np.uint8, 1),
          ("fit", np.float32, 4), ("radius", np.int32, 1), ("neighbours", np.int16, 18), ("reliable", np.uint8, 1)]


def camera_bytes(K, R, t, w, h, full_w, full_h, dmin, dmax):
    """struct Camera (main.h:50-59) at the stage's scale (K scaled as DPE.cpp:804-817, float arithmetic)."""
    K = np.array(K, np.float32).copy()
    R, t = np.asarray(R, np.float32), np.asarray(t, np.float32)
    if (w, h) != (full_w, full_h):
        sx, sy = np.float32(w) / np.float32(full_w), np.float32(h) / np.float32(full_h)
        K[0, 0] *= sx; K[0, 2] *= sx; K[1, 1] *= sy; K[1, 2] *= sy
    c = -(R.astype(np.float64).T @ t.astype(np.float64)).astype(np.float32)
    return K.tobytes() + R.tobytes() + t.tobytes() + c.tobytes() + struct.pack("<iiff", h, w, dmin, dmax)


def run_probe(tag, imgs, cams, full_wh, drange, p, planes, state, selected, src_depths, edge, edge_low, label):
    H, W = imgs[0].shape
    N = len(imgs)
    lw, lh = edge_low.shape[1], edge_low.shape[0]
    blob = struct.pack("<11i", W, H, N, lw, lh, p.state, p.geom_consistency, p.use_apd, p.max_iterations, p.weak_peak_radius, p.rotate_time)
    blob += struct.pack("<3f", p.ransac_threshold, np.float32(drange[0]) * np.float32(0.6), np.float32(drange[1]) * np.float32(1.2))
    blob += np.stack(imgs).astype(np.float32).tobytes()
    if p.geom_consistency:
        blob += np.stack([np.zeros((H, W), np.float32)] + [d.astype(np.float32) for d in src_depths]).tobytes()
    for (K, R, t) in cams:
        blob += camera_bytes(K, R, t, W, H, full_wh[0], full_wh[1], drange[0], drange[1])
    blob += planes.astype(np.float32).tobytes() + state.astype(np.uint8).tobytes() + selected.astype(np.uint32).tobytes()
    blob += edge.astype(np.uint8).tobytes() + edge_low.astype(np.uint8).tobytes() + label.astype(np.int32).tobytes()
    fin, fout = Path(f"/tmp/stage_{tag}_in.bin"), Path(f"/tmp/stage_{tag}_out.bin")
    fin.write_bytes(blob)
    subprocess.check_call([str(ROOT / "oracle" / "_ref" / "ref_stage_probe"), str(fin), str(fout)])
    raw = fout.read_bytes()
    P = W * H
    out, off = {}, 0
    while off < len(raw):
        (step,) = struct.unpack_from("<i", raw, off); off += 4
        rec = {}
        for name, dt, ch in FIELDS:
            n = P * ch
            a = np.frombuffer(raw, dt, n, off).copy(); off += n * np.dtype(dt).itemsize
            rec[name] = a.reshape((H, W) if ch == 1 else (H, W, ch) if name != "neighbours" else (H, W, 9, 2))
        out[step] = rec
    return out


def step_compare(planes, sel, ref, costs=None):
    """per-pixel state after one step (camera-coordinate plane hypotheses, selected views) against the
    reference kernels' dump of the same step"""
    dn = np.abs(planes[..., :3] - ref["planes"][..., :3]).max(-1)
    dd = np.abs(planes[..., 3] - ref["planes"][..., 3]) / np.maximum(np.abs(ref["planes"][..., 3]), 1e-9)
    out = {"plane_identical": float(((dn < 1e-4) & (dd < 1e-4)).mean()), "plane_bitwise": float((planes == ref["planes"]).all(-1).mean()),
           "selected_equal": float((sel == ref["selected"]).mean())}
    if costs is not None:
        out["cost_bitwise"] = float(((costs == ref["costs"]) | (np.isnan(costs) & np.isnan(ref["costs"]))).mean())
    return out


def final_compare(fin, ref_fin, drange):
    """final maps of a stage (dpe_get_maps) against the reference kernels' last dump + the host tail of
    ProcessProblem (main.cpp:427-437: depths outside the PatchMatch range -> 0 / UNKNOWN)."""
    rd = ref_fin["planes"][..., 3].copy()
    rst = ref_fin["state"].copy()
    dmin, dmax = np.float32(drange[0]) * np.float32(0.6), np.float32(drange[1]) * np.float32(1.2)
    oob = (rd < dmin) | (rd > dmax)
    rd[oob] = 0; rst[oob] = 2
    both = (fin["depth"] > 0) & (rd > 0)
    rel_d = np.abs(fin["depth"] - rd) / np.maximum(rd, 1e-9)
    dn = np.abs(fin["normal"] - ref_fin["planes"][..., :3]).max(-1)
    ang = np.degrees(np.arccos(np.clip((fin["normal"] * ref_fin["planes"][..., :3]).sum(-1), -1, 1)))
    return {"final_state_equal": float((fin["state"] == rst).mean()),
            "final_depth_1e-3": float((rel_d[both] < 1e-3).mean()), "final_depth_1pct": float((rel_d[both] < 1e-2).mean()),
            "final_normal_identical": float((dn[both] < 1e-4).mean()), "final_normal_1deg": float((ang[both] < 1).mean())}


def all_stages(report, grays, cams, drs, pairs, prep, sizes, sched, v):
    """Every stage of the schedule for view v: the reference's kernels (probe) run on the maps THIS implementation
    holds after the previous stage, and the stage's final maps are compared.  Tells which stage type still
    differs (FIRST_INIT / coarse REFINE_ITER with geometric consistency / REFINE_INIT across the scale change /
    fine REFINE_ITER with the weak path)."""
    import json
    W, H = sizes[-1]
    ids = [v] + list(pairs[v])
    ctx = capi.Context(0)
    capi.upload_scene(ctx, grays, cams, drs, pairs, 2)
    for vv in range(len(grays)):
        for kk in range(2):
            ctx.set_prep(vv, kk, *prep[vv][kk])
    ctx.set_reference_race(int(os.environ.get("DPE_STAGE_DIFF_RACE", "1")))   # 1 live, 2 pre-sweep copy
    ctx.set_profile(len(grays))
    out = {}
    prev_k = None
    maps = None
    for si, (k, p) in enumerate(sched):
        w, h = sizes[k]
        imgs = [grays[i].astype(np.float32) if (w, h) == (W, H) else hostsim.resize_linear(grays[i].astype(np.float32), w, h) for i in ids]
        if si == 0:
            planes = np.zeros((h, w, 4), np.float32); state = np.ones((h, w), np.uint8); sel = np.zeros((h, w), np.uint32)
        else:
            pm = maps[v]
            planes = np.concatenate([pm["normal"], pm["depth"][..., None]], -1).astype(np.float32)
            state, sel = pm["state"], pm["selected"]
            if prev_k != k:      # RescaleMatToTargetSize (DPE.cpp:1147-1168): nearest neighbour, factors swapped
                ph, pw_ = state.shape
                sx, sy = np.float32(w) / np.float32(pw_), np.float32(h) / np.float32(ph)
                oy = (np.arange(h, dtype=np.float32) / sx).astype(np.int64); ox = (np.arange(w, dtype=np.float32) / sy).astype(np.int64)
                ok = (oy[:, None] < ph) & (ox[None, :] < pw_)
                oyc, oxc = np.minimum(oy, ph - 1), np.minimum(ox, pw_ - 1)
                planes = np.where(ok[..., None], planes[oyc][:, oxc], 0).astype(np.float32)
                state = np.where(ok, state[oyc][:, oxc], 1).astype(np.uint8)
                sel = np.where(ok, sel[oyc][:, oxc], 0).astype(np.uint32)
        src_d = [maps[i]["depth"] for i in pairs[v]] if p.geom_consistency else None
        if p.use_apd:
            edge, edge_low, label = prep[v][k][0], prep[v][0][0], prep[v][k][1]
        else:
            edge = np.zeros((h, w), np.uint8); edge_low = np.zeros(sizes[0][::-1], np.uint8); label = np.full((h, w), -1, np.int32)
        dumps = run_probe(f"all{si}", imgs, [cams[i] for i in ids], (W, H), drs[v], p, planes, state, sel, src_d, edge, edge_low, label)
        steps = {}
        if si > 0:               # per step (a truncated stage leaves the carried maps and the atlas untouched)
            for step in ((1, 2, 5, 8) if not p.use_apd else (1, 2, 4, 7, 10)):
                ctx.debug_stop_after(step)
                ctx.run_stage(k, p, SEED)
                steps[f"step{step}"] = step_compare(ctx.debug_read(7, (h, w, 4), np.float32), ctx.debug_read(8, (h, w), np.uint32), dumps[step],
                                                    ctx.debug_read(3, (h, w), np.float32))
            ctx.debug_stop_after(-1)
        ctx.run_stage(k, p, SEED); ctx.stage_commit()
        maps = [ctx.get_maps(i, k) for i in range(len(grays))]
        prev_k = k
        r = final_compare(maps[v], dumps[11], drs[v])
        r.update(steps)
        if os.environ.get("DPE_STAGE_DIFF_REFREF") and si > 0:
            # the reference against itself: the same probe once more on the same inputs (its direction-4 race)
            d2 = run_probe(f"all{si}b", imgs, [cams[i] for i in ids], (W, H), drs[v], p, planes, state, sel, src_d, edge, edge_low, label)
            rr = {f"step{st_}_plane_bitwise": float((d2[st_]["planes"].view(np.uint32) == dumps[st_]["planes"].view(np.uint32)).all(-1).mean())
                  for st_ in ((1, 2, 5, 8) if not p.use_apd else (1, 2, 4, 7, 10))}
            ok2 = (d2[11]["planes"][..., 3] > 0) & (dumps[11]["planes"][..., 3] > 0)
            dn2 = np.abs(d2[11]["planes"][..., :3] - dumps[11]["planes"][..., :3]).max(-1)
            rr["final_normal_identical"] = float((dn2[ok2] < 1e-4).mean())
            rr["final_state_equal"] = float((d2[11]["state"] == dumps[11]["state"]).mean())
            r["ref_vs_ref"] = rr
        r["selected_equal"] = float((maps[v]["selected"] == dumps[11]["selected"]).mean())
        r["weak_frac"] = float((maps[v]["state"] == 0).mean())
        out[f"stage{si}"] = r
        print(f"stage {si} (scale {k}, state {p.state}, geom {p.geom_consistency}, apd {p.use_apd}):",
              json.dumps({a: (round(b, 4) if not isinstance(b, dict) else {c: round(d, 4) for c, d in b.items()}) for a, b in r.items()}), flush=True)
    ctx.close()
    report["gpu_vs_ref_every_stage"] = out


def main():
    spec, grays, cams, drs, pairs, gt = small_scene("c4", 0.05, 5)      # 151 x 101, low-texture planes
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
    v = 0
    ids = [v] + list(pairs[v])
    sched = capi.stage_schedule(2)
    report = {}
    # ---- case "first": stage 0 at the coarse scale
    k, p = sched[0]
    cw, ch = sizes[0]
    imgs0 = [hostsim.resize_linear(grays[i].astype(np.float32), cw, ch) for i in ids]
    zero_e = np.zeros((ch, cw), np.uint8)
    dumps = run_probe("first", imgs0, [cams[i] for i in ids], (W, H), drs[v], p, np.zeros((ch, cw, 4), np.float32),
                      np.ones((ch, cw), np.uint8), np.zeros((ch, cw), np.uint32), None, zero_e, zero_e, np.full((ch, cw), -1, np.int32))
    rec = dict(images=np.stack(imgs0), K=np.stack([cams[i][0] for i in ids]), R=np.stack([cams[i][1] for i in ids]),
               t=np.stack([cams[i][2] for i in ids]), drange=np.array(drs[v], np.float32), full_wh=np.array([W, H]))
    for s in (1, 2, 8, 11):
        for name in ("planes", "costs", "selected", "state"):
            rec[f"s{s}_{name}"] = dumps[s][name]
    np.savez_compressed(ROOT / "tests" / "golden" / "ref_stage_first.npz", **rec)
    print("first: steps", sorted(dumps))
    import json
    # the same stage on the GPU (this implementation) in the three cost arithmetics' reference variants
    for arith in (1, 2):
        c0 = capi.Context(0)
        capi.upload_scene(c0, grays, cams, drs, pairs, 2)
        c0.set_profile(len(grays))
        c0.set_cost_arithmetic(arith)
        c0.run_stage(*sched[0], SEED); c0.stage_commit()
        fin = c0.get_maps(v, 0)
        c0.close()
        r = final_compare(fin, dumps[11], drs[v])
        for step in (1, 2, 5, 8):
            c0 = capi.Context(0)
            capi.upload_scene(c0, grays, cams, drs, pairs, 2, active=(v, 1))
            c0.set_cost_arithmetic(arith)
            c0.debug_stop_after(step)
            c0.run_stage(*sched[0], SEED)
            r[f"step{step}"] = step_compare(c0.debug_read(7, (ch, cw, 4), np.float32), c0.debug_read(8, (ch, cw), np.uint32), dumps[step], c0.debug_read(3, (ch, cw), np.float32))
            c0.close()
        report[f"gpu_vs_ref_stage0_arith{arith}"] = r
        print(f"GPU vs reference kernels, stage 0, view {v}, arithmetic={arith}:", json.dumps(r))
    # ---- case "weak": stage 6, inputs = this implementation's maps after stages 0..5 (GPU)
    ctx = capi.Context(0)
    capi.upload_scene(ctx, grays, cams, drs, pairs, 2)
    for vv in range(len(grays)):
        for kk in range(2):
            ctx.set_prep(vv, kk, *prep[vv][kk])
    for si in range(6):
        ctx.run_stage(*sched[si], SEED); ctx.stage_commit()
    maps = [ctx.get_maps(i, 1) for i in range(len(grays))]
    # the weak case uses the LAST view: the GPU run of stage 6 below keeps its scratch arrays readable
    v = len(grays) - 1
    ids = [v] + list(pairs[v])
    k, p = sched[6]
    pv = maps[v]
    planes = np.concatenate([pv["normal"], pv["depth"][..., None]], -1).astype(np.float32)
    src_d = [maps[i]["depth"] for i in pairs[v]]
    imgs1 = [grays[i].astype(np.float32) for i in ids]
    dumps = run_probe("weak", imgs1, [cams[i] for i in ids], (W, H), drs[v], p, planes, pv["state"], pv["selected"], src_d,
                      prep[v][1][0], prep[v][0][0], prep[v][1][1])
    rec = dict(images=np.stack([grays[i] for i in ids]), K=np.stack([cams[i][0] for i in ids]), R=np.stack([cams[i][1] for i in ids]),
               t=np.stack([cams[i][2] for i in ids]), drange=np.array(drs[v], np.float32), full_wh=np.array([W, H]),
               prev_planes=planes, prev_state=pv["state"], prev_selected=pv["selected"], src_depths=np.stack(src_d),
               edge=prep[v][1][0], edge_low=prep[v][0][0], label=prep[v][1][1])
    want = {0: ("state", "neighbours", "reliable"), 1: ("planes", "costs", "selected"), 2: ("planes", "costs", "selected"),
            3: ("fit", "radius"), 4: ("planes", "costs", "selected"), 10: ("planes", "costs", "selected"), 11: ("planes", "state", "selected")}
    for s, names in want.items():
        for name in names:
            rec[f"s{s}_{name}"] = dumps[s][name]
    np.savez_compressed(ROOT / "tests" / "golden" / "ref_stage_weak.npz", **rec)
    nweak = int((pv["state"] == 0).sum())
    print("weak: steps", sorted(dumps), "weak pixels in", nweak, "of", pv["state"].size)
    # ---- the reference against itself: the same probe once more (its direction-4 race is the only difference)
    again = run_probe("weak2", imgs1, [cams[i] for i in ids], (W, H), drs[v], p, planes, pv["state"], pv["selected"], src_d,
                      prep[v][1][0], prep[v][0][0], prep[v][1][1])
    wm0 = pv["state"] == 0
    rr = {}
    for s_ in (2, 4, 10, 11):
        dn_ = np.abs(again[s_]["planes"][..., :3] - dumps[s_]["planes"][..., :3]).max(-1)
        rr[f"step{s_}_normal_identical"] = float((dn_ < 1e-4).mean())
    rr["anchor_sets_equal_incl_order"] = float((again[0]["neighbours"][wm0] == dumps[0]["neighbours"][wm0]).all(-1).all(-1).mean())
    rr["final_state_equal"] = float((again[11]["state"] == dumps[11]["state"]).mean())
    a11, b11 = again[11]["planes"], dumps[11]["planes"]
    ang_ = np.degrees(np.arccos(np.clip((a11[..., :3] * b11[..., :3]).sum(-1), -1, 1)))
    ok_ = (a11[..., 3] > 0) & (b11[..., 3] > 0)
    rr["final_normal_1deg"] = float((ang_[ok_] < 1).mean())
    rr["final_depth_1pct"] = float((np.abs(a11[..., 3] - b11[..., 3])[ok_] / b11[..., 3][ok_] < 1e-2).mean())
    report["ref_vs_ref_stage6"] = rr
    print("reference kernels vs themselves (second run), stage 6:", json.dumps(rr))
    # ---- the same stage on the GPU (this implementation), all views on the first stream so that the last
    # view's scratch arrays survive; compared with the reference's dumps
    for race, arith in ((0, 1), (1, 1), (1, 2)):
        c2 = capi.Context(0)
        capi.upload_scene(c2, grays, cams, drs, pairs, 2)
        for vv in range(len(grays)):
            for kk in range(2):
                c2.set_prep(vv, kk, *prep[vv][kk])
        for si in range(6):
            c2.run_stage(*sched[si], SEED); c2.stage_commit()
        c2.set_profile(len(grays))
        c2.set_reference_race(race)
        c2.set_cost_arithmetic(arith)
        c2.run_stage(*sched[6], SEED); c2.stage_commit()
        nb = c2.debug_read(0, (H, W, 9, 2), np.int16)
        fit = c2.debug_read(1, (H, W, 4), np.float32)
        rad = c2.debug_read(2, (H, W), np.int32)
        rel = c2.debug_read(4, (H, W), np.uint8)
        fin = c2.get_maps(v, 1)
        c2.close()
        wm = pv["state"] == 0
        r = {}
        eq = (nb[wm] == dumps[0]["neighbours"][wm]).all(-1)
        r["anchors_equal"] = float(eq.mean()); r["anchor_sets_equal"] = float(eq.all(-1).mean())
        r["reliable_equal"] = float((rel[wm] == dumps[0]["reliable"][wm]).mean())
        r["radius_equal_weak"] = float((rad[wm] == dumps[9]["radius"][wm]).mean())
        dn = np.abs(fit[..., :3] - dumps[9]["fit"][..., :3]).max(-1)
        r["fit_normal_equal_weak"] = float((dn[wm] < 1e-4).mean())
        r.update(final_compare(fin, dumps[11], drs[v]))
        if race == 1:
            for step in (1, 2, 4, 7, 10):
                c3 = capi.Context(0)
                capi.upload_scene(c3, grays, cams, drs, pairs, 2)
                for vv in range(len(grays)):
                    for kk in range(2):
                        c3.set_prep(vv, kk, *prep[vv][kk])
                for si in range(6):
                    c3.run_stage(*sched[si], SEED); c3.stage_commit()
                c3.set_profile(len(grays))
                c3.set_reference_race(race)
                c3.set_cost_arithmetic(arith)
                c3.debug_stop_after(step)
                c3.run_stage(*sched[6], SEED)
                sc_ = step_compare(c3.debug_read(7, (H, W, 4), np.float32), c3.debug_read(8, (H, W), np.uint32), dumps[step], c3.debug_read(3, (H, W), np.float32))
                pl3 = c3.debug_read(7, (H, W, 4), np.float32)
                dn3 = np.abs(pl3[..., :3] - dumps[step]["planes"][..., :3]).max(-1)
                sc_["plane_identical_weakpx"] = float((dn3[wm] < 1e-4).mean())
                r[f"step{step}"] = sc_
                c3.close()
        both = (fin["depth"] > 0) & (dumps[11]["planes"][..., 3] > 0)
        rel_d = np.abs(fin["depth"] - dumps[11]["planes"][..., 3]) / np.maximum(dumps[11]["planes"][..., 3], 1e-9)
        r["final_weakpx_depth_1pct"] = float((rel_d[both & wm] < 1e-2).mean())
        key = f"gpu_vs_ref_stage6_race{race}" + ("_exact" if arith == 2 else "")
        report[key] = r
        print(f"GPU vs reference kernels, stage 6, view {v}, ref_race={race}, arithmetic={arith}:", json.dumps(r))
    all_stages(report, grays, cams, drs, pairs, prep, sizes, sched, len(grays) - 1)
    (ROOT / "gpurun_out" / "stage_diff.json").write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
