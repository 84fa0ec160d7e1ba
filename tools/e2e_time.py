"""Times DPE_MVS.dpe_mvs() on the bench scene under different view orders / GPU counts."""
import json, os, shutil, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT))
from bench import ensure_scene
import DPE_MVS
folder = ensure_scene("c2", None, "c2")
out = {}
for order in ("sequential", "parallel"):
    os.environ["DPE_VIEW_ORDER"] = order
    tj = folder / "timing.json"
    os.environ["DPE_TIMING_JSON"] = str(tj)
    shutil.rmtree(folder / "DPE", ignore_errors=True)
    t0 = time.perf_counter()
    DPE_MVS.dpe_mvs(str(folder), 0, False, False, False, True, False, False, False)
    out[order] = dict(seconds=time.perf_counter() - t0, breakdown=json.loads(tj.read_text()))
    print(order, out[order], flush=True)
(ROOT / "gpurun_out" / "e2e_time.json").write_text(json.dumps(out, indent=1))
