"""Attributes the warp-state samples of an .ncu-rep to CUDA source lines by joining ncu's SASS
page with nvdisasm's line info of the cubin built from the same object (instruction order).
usage: python tools/ncu_lines.py <rep> <object.o> <kernel mangled name> [top]"""
import csv, io, re, subprocess, sys, tempfile, os, collections
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sec = dis.split(".text." + kern + ":", 1)[1]
sec = sec.split("//--------------------- .", 1)[0]
lines = []   # per instruction: (file, line)
cur = ("?", 0)
for l in sec.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
samples = []
for r in rows[2:]:
    try:
        samples.append((int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]] or 0), r[ix["Source"]].strip()))
    except Exception:
        pass
print(f"instructions: nvdisasm {len(lines)}  ncu {len(samples)}")
n = min(len(lines), len(samples))
agg = collections.Counter(); ex = collections.Counter()
for i in range(n):
    agg[lines[i]] += samples[i][0]; ex[lines[i]] += samples[i][1]
tot = sum(agg.values())
byfile = collections.Counter()
for (f, l), s in agg.items():
    byfile[f] += s
print("by file:", {k: f"{100.0*v/tot:.1f}%" for k, v in byfile.most_common()})
srcs = {}
for (f, l), s in agg.most_common(top):
    if f not in srcs:
        for root in ("dpe-mvs_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                srcs[f] = open(p).read().split("\n"); break
        else:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{s:7d} {100.0*s/tot:5.1f}%  x{ex[(f,l)]:>11d}  {f}:{l}  {text}")
