"""Every stage of the schedule for one view of a full-size synthetic scene: the reference's own kernels
(oracle/_ref/ref_stage_probe) run on the maps THIS implementation holds after the previous stage, state compared
bit for bit after each launch group (oracle/make_stage_golden.py: all_stages).  The committed fixtures replay a
151 x 101 scene, where the edge-adaptive sampling never leaves its minimum step; this runs the same comparison at the
sizes where it does.  GPU box.   usage: stage_diff_scene.py <config> <scale> <views> <view> [race mode 1|2]"""
import ctypes as C
import json
import os
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200")); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
config, scale, n_views, v = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
if len(sys.argv) > 5:
    os.environ["DPE_STAGE_DIFF_RACE"] = sys.argv[5]
import capi, simpipe  # noqa: E402
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
report = {"scene": config, "size": [W, H], "views": len(grays), "view": v, "race_mode": os.environ.get("DPE_STAGE_DIFF_RACE", "1")}
msg.all_stages(report, grays, cams, drs, pairs, prep, sizes, capi.stage_schedule(2), v)
out = ROOT / "gpurun_out" / f"stage_diff_{config}_{W}x{H}_race{report['race_mode']}.json"
out.write_text(json.dumps(report, indent=1))
