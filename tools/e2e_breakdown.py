"""Times DPE_MVS.dpe_mvs() on the bench scene and prints the pipeline's own breakdown (DPE_TIMING_JSON);
   usage: e2e_breakdown.py [gpus=1] [repeats=2] [config=c2]"""
import json, os, shutil, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT))
from bench import ensure_scene
import DPE_MVS
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = sys.argv[3] if len(sys.argv) > 3 else "c2"
folder = ensure_scene(cfg, None, cfg)
if G > 1:
    os.environ["DPE_GPUS"] = ",".join(str(i) for i in range(G))
tj = folder / "timing.json"
os.environ["DPE_TIMING_JSON"] = str(tj)
out = []
for r in range(reps):
    shutil.rmtree(folder / "DPE", ignore_errors=True)
    t0 = time.perf_counter()
    DPE_MVS.dpe_mvs(str(folder), 0 if G == 1 else -1, False, False, False, True, False, False, False)
    dt = time.perf_counter() - t0
    out.append(dict(seconds=dt, breakdown=json.loads(tj.read_text())))
    print(json.dumps(out[-1]), flush=True)
(ROOT / "gpurun_out" / f"e2e_breakdown_{cfg}_g{G}.json").write_text(json.dumps(out, indent=1))
