"""Texture-pipe study on the B200: filtered-fetch rate of the NCC tap pattern as a function of
texel format, warp lane layout, reference->source map and occupancy.  Writes gpurun_out/tex_study.json."""
import json, sys, math
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "dpe-mvs_b200"))
import capi
ctx = capi.Context(0)
out = []
def rot(deg, s=1.0):
    c, si = math.cos(math.radians(deg)) * s, math.sin(math.radians(deg)) * s
    return (c, -si, si, c)
maps = {"identity": (1, 0, 0, 1), "rot10": rot(10), "rot30": rot(30), "rot90": rot(90), "scale0.7": rot(0, 0.7), "scale1.4": rot(0, 1.4),
        "shear": (1.0, 0.3, 0.1, 0.9)}
for fmt, fname in ((0, "f32"), (1, "f16"), (2, "u8")):
    for layout, lname in ((0, "row32"), (1, "zigzag"), (2, "8x4"), (3, "colour8x8"), (4, "16x2")):
        for mname, m in maps.items():
            r = max(ctx.probe_tex_pattern(fmt, layout, m) for _ in range(2))
            out.append(dict(fmt=fname, layout=lname, map=mname, threads=128, bps=4, gtaps=r / 1e9))
            print(out[-1], flush=True)
# occupancy: resident warps per SM
for (threads, bps) in ((128, 2), (128, 4), (128, 6), (128, 8), (256, 8), (128, 16)):
    for fmt in (0, 2):
        r = max(ctx.probe_tex_pattern(fmt, 1, maps["rot10"], threads=threads, blocks_per_sm=bps) for _ in range(2))
        out.append(dict(fmt=fmt, layout="zigzag", map="rot10", threads=threads, bps=bps, gtaps=r / 1e9))
        print(out[-1], flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "tex_study.json").write_text(json.dumps(out, indent=1))
