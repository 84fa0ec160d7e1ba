"""Static look at texture-level parallelism in the SASS of one kernel: for every TEX, how many
earlier TEX results are still unconsumed when it issues (straight-line approximation).
usage: python tools/sass_tex_mlp.py <object.o> <mangled kernel name substring>"""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            ins.append(m.group(2).strip())
    pending = {}   # dest reg -> index of TEX
    hist = []
    for i, s in enumerate(ins):
        body = re.sub(r"^@!?U?P\d+\s+", "", s)
        op = body.split()[0]
        regs = re.findall(r"\bR(\d+)\b", body)
        if op.startswith("TEX"):
            # operands: TEX.LL RZ, Rdst, Rcoord, Rlod...
            ops = [o.strip() for o in body.split(None, 1)[1].split(",")]
            dst = ops[1]
            srcs = re.findall(r"\bR(\d+)\b", ",".join(ops[2:]))
            for r in srcs:
                pending.pop("R" + r, None)
            hist.append(len(pending))
            pending[dst] = i
        else:
            ops = body.split(None, 1)[1] if " " in body else ""
            parts = [o.strip() for o in ops.split(",")]
            for r in re.findall(r"\bR(\d+)\b", ",".join(parts[1:])):
                pending.pop("R" + r, None)
            # a write to a pending dest also retires it
            if parts and re.match(r"R\d+$", parts[0]):
                pending.pop(parts[0], None)
            if op in ("BRA", "CALL.REL.NOINC", "RET.REL.NODEC", "EXIT", "BSYNC"):
                pass
    if hist:
        print(f"{name[:70]:70s} TEX={len(hist):4d}  in-flight at issue: mean={sum(hist)/len(hist):5.1f} max={max(hist):3d}  hist={hist[:40]}")
