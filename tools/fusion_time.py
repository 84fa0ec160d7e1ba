"""Times DPE_MVS.dpe_mvs(fusion=True) on the bench scene (49 x 1600x1200): device fusion + DPE.ply."""
import json, os, shutil, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT))
from bench import ensure_scene
import DPE_MVS
folder = ensure_scene("c2", None, "c2")
tj = folder / "timing.json"
os.environ["DPE_TIMING_JSON"] = str(tj)
shutil.rmtree(folder / "DPE", ignore_errors=True)
t0 = time.perf_counter()
DPE_MVS.dpe_mvs(str(folder), 0, False, True, False, True, False, False, False)
dt = time.perf_counter() - t0
b = json.loads(tj.read_text())
ply = folder / "DPE" / "DPE.ply"
head = ply.read_bytes()[:200].decode(errors="ignore")
n = int([l for l in head.split("\n") if l.startswith("element vertex")][0].split()[-1])
out = dict(seconds=dt, breakdown=b, points=n, ply_bytes=ply.stat().st_size)
print(json.dumps(out))
(ROOT / "gpurun_out" / "fusion_time.json").write_text(json.dumps(out, indent=1))
