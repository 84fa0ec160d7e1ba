"""GPU cost kernel vs the reference's own device outputs (tests/golden/ref_probe_c1.npz) in the three cost
arithmetics."""
import sys, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200"))
import capi
fx = np.load(ROOT / "tests" / "golden" / "ref_probe_c1.npz")
imgs = [fx["images"][i] for i in range(4)]
cams = [(fx["K"][i], fx["R"][i], fx["t"][i]) for i in range(4)]
out = {}
for mode, name in ((1, "reference"), (2, "reference_exact"), (0, "centred")):
    ctx = capi.Context(0)
    H, W = imgs[0].shape
    ctx.scene_begin(4, W, H, 1)
    for v in range(4):
        ctx.set_view(v, imgs[v], *cams[v], 1.0, 10.0)
    ctx.set_pairs(0, [1, 2, 3])
    ctx.commit()
    ctx.set_cost_arithmetic(mode)
    got = ctx.cost_eval(0, 0, fx["xy"], fx["planes"], 3, mode=0)
    ref = fx["ref_ncc"]
    both = (got < 2.0) & (ref < 2.0)
    d = np.abs(got - ref)[both]
    out[name] = dict(n=int(both.sum()), same_invalid=float(((got >= 2.0) == (ref >= 2.0)).mean()), median=float(np.median(d)),
                     p90=float(np.percentile(d, 90)), p99=float(np.percentile(d, 99)), max=float(d.max()), exact=float((d == 0).mean()),
                     below_1e6=float((d < 1e-6).mean()), below_1e5=float((d < 1e-5).mean()))
    print(name, json.dumps(out[name]))
    if mode == 2:
        bad = np.argwhere(both & (got != ref))
        print("differing evaluations (pixel index, view): ", len(bad))
        for (i, v) in bad[:60]:
            print(f"  i={i} v={v} xy={fx['xy'][i].tolist()} plane={fx['planes'][i].tolist()} got={got[i, v]:.9g} ref={ref[i, v]:.9g} d={got[i, v] - ref[i, v]:.3g}")
        bad_px = sorted(set(int(b[0]) for b in bad))
        print("pixels with a difference:", len(bad_px), "of", len(fx['xy']), "; views:", np.bincount(bad[:, 1], minlength=3).tolist())
    # geometric consistency against the reference's ComputeGeomConsistencyCost outputs (golden source depths
    # written into the atlas)
    import torch
    ctx.run_stage(*capi.stage_schedule(1)[0], 5)
    ptr, slot, total = ctx.stage_atlas()

    class _Raw:
        __cuda_array_interface__ = {"shape": (4, H, W), "typestr": "<f4", "data": (ptr, False), "version": 2}
    torch.as_tensor(_Raw(), device="cuda:0").copy_(torch.from_numpy(np.ascontiguousarray(fx["depths"], np.float32)))
    torch.cuda.synchronize()
    ctx.stage_commit()
    gg = ctx.geom_eval(0, 0, fx["xy"], fx["planes"], 3)
    rg = fx["ref_geom"]
    dg = np.abs(gg - rg)
    near = dg < 1e-2
    out[name + "_geom"] = dict(same_3=float(((gg == 3.0) == (rg == 3.0)).mean()), within_1e2=float(near.mean()), exact=float((dg == 0).mean()),
                               median=float(np.median(dg[near])), p90=float(np.percentile(dg[near], 90)), p99=float(np.percentile(dg[near], 99)))
    print(name + "_geom", json.dumps(out[name + "_geom"]))
    ctx.close()
(ROOT / "gpurun_out" / "cost_diff.json").write_text(json.dumps(out, indent=1))
