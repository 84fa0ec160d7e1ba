"""Runs DPE_MVS.dpe_mvs() on one of the BASELINE.json scene shapes (optionally reduced) and reports timing and
accuracy against the analytic ground truth.  usage: run_config.py <c1|c2|c4|c5> [--views N] [--scale S] [--fusion] [--gpus K]"""
import argparse, json, os, shutil, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200"))
import synth, DPE_MVS
ap = argparse.ArgumentParser()
ap.add_argument("config"); ap.add_argument("--views", type=int, default=None); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--fusion", action="store_true"); ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--sharded-fusion", action="store_true"); ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--no-sidecar", action="store_true", help="skip the .gray sidecars (only the reference build's imread stand-in needs them)")
ap.add_argument("--no-normal", action="store_true", help="do not write normal.npy (25 MB per 1920x1080 view)")
ap.add_argument("--keep-scene", action="store_true", help="reuse the scene folder of an earlier invocation")
args = ap.parse_args()
tag = f"{args.config}_v{args.views}_s{args.scale}_g{args.gpus}" + ("_shardedfusion" if args.sharded_fusion else "")
folder = Path("/tmp") / f"cfg_{args.config}_v{args.views}_s{args.scale}"
spec = synth.make_scene(args.config, scale=args.scale, n_views=args.views)
gt_every = max(1, spec.n_views // 8)
t_gen = 0.0
if not (args.keep_scene and (folder / "pair.txt").exists()):
    shutil.rmtree(folder, ignore_errors=True)
    t0 = time.time(); synth.write_scene(spec, folder, save_gt="depth", sidecar=not args.no_sidecar, gt_every=gt_every); t_gen = time.time() - t0
if args.sharded_fusion:
    os.environ["DPE_FUSION_SHARDED"] = "1"
if args.gpus > 1:
    os.environ["DPE_GPUS"] = ",".join(str(i) for i in range(args.gpus))
tj = folder / "timing.json"; os.environ["DPE_TIMING_JSON"] = str(tj)
weak = args.config == "c4"
runs = []
for rep in range(args.repeat):          # the first call of a process also brings up CUDA contexts and the NCCL communicator
    shutil.rmtree(folder / "DPE", ignore_errors=True)
    t0 = time.perf_counter()
    DPE_MVS.dpe_mvs(str(folder), 0 if args.gpus == 1 else -1, False, args.fusion, False, True, not args.no_normal, weak, weak)
    runs.append((time.perf_counter() - t0, json.loads(tj.read_text())))
dt = runs[-1][0]
out = dict(config=args.config, views=spec.n_views, width=spec.width, height=spec.height, n_src=spec.n_src, gpus=args.gpus,
           fusion=args.fusion, sharded_fusion=args.sharded_fusion, scene_gen_s=t_gen, seconds=dt, depth_maps_per_s=spec.n_views / dt,
           seconds_all_calls=[r[0] for r in runs], breakdown=runs[-1][1], first_call_breakdown=runs[0][1])
acc = []
for v in range(0, spec.n_views, gt_every):
    d = np.load(folder / "DPE" / f"{v:08d}" / "depth.npy"); g = np.load(folder / "gt" / f"{v:08d}_depth.npy")
    m = (g > 0) & (d > 0); rel = np.abs(d - g) / np.maximum(g, 1e-9)
    acc.append(dict(view=v, cover=float((d > 0).mean()), depth_1pct=float((rel[m] < 0.01).mean())))
out["accuracy"] = acc
if args.fusion:
    ply = folder / "DPE" / "DPE.ply"
    head = ply.read_bytes()[:200].decode(errors="ignore")
    out["ply_points"] = int([l for l in head.split("\n") if l.startswith("element vertex")][0].split()[-1])
print(json.dumps(out))
(ROOT / "gpurun_out" / f"cfg_{tag}.json").write_text(json.dumps(out, indent=1))
