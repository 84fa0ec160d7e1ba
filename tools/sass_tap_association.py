"""Which product of a tap coordinate does each inlined NCC copy of a reference kernel round on its own?
Reads cuobjdump -sass text of one kernel (no GPU needed) and classifies every source-tap texture fetch:
A = the x products (hoisted out of the inner tap loop, y products fused onto them), B = the y products.

  cuobjdump -sass oracle/_ref/ref_stage_probe | awk '/Function : <kernel>/{f=1;next} /Function :/{f=0} f' > k.sass
  python tools/sass_tap_association.py k.sass

Result for the reference built with nvcc 12.9 for sm_100 (oracle/Makefile): Black/RedPixelUpdateStrong have 12
copies — A A B B B B B B B A A A in code order = edge-mode sites, the eight ACMM slots (up_far, down_far, left_far,
right_far, up_near, down_near, left_near = B; right_near = A), re-score, refinement; every copy in
RandomInitialization, DepthToWeak, LocalRefine and Black/RedPixelUpdateWeak is A.
"""
import re,sys
ins=[]
for l in open(sys.argv[1]):
    m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l)
    if m: ins.append((m.group(1),m.group(2).strip()))
addr={a:i for i,(a,t) in enumerate(ins)}
tex=[i for i,(a,t) in enumerate(ins) if 'TEX' in t]
# src-tap TEX = those whose coordinates come from FFMA ...,0.5 ; find for each such TEX the inner loop start (next backward BRA target after it)
res=[]
for i in tex:
    # coordinates defined by FFMA x, x, x, 0.5 within 12 instr before
    win=ins[max(0,i-14):i]
    if sum(1 for a,t in win if t.startswith('FFMA') and t.endswith('0.5'))<2: continue
    # inner loop start: first backward BRA after i (skip BRA.U.ANY)
    start=None
    for j in range(i,min(len(ins),i+400)):
        t=ins[j][1]
        m=re.search(r'BRA 0x([0-9a-f]+)',t)
        if m and 'ANY' not in t:
            tgt=int(m.group(1),16)
            if tgt < int(ins[j][0],16) and tgt <= int(ins[i][0],16):
                start=tgt; break
    # the three FMULs feeding the FFMAs: find FFMA (non-0.5) in window of 40 before, their addend regs defined by FMUL: is that FMUL inside the inner loop?
    w=ins[max(0,i-45):i]
    ffma=[(a,t) for a,t in w if t.startswith('FFMA') and not t.endswith('0.5')]
    inside=0;outside=0
    for a,t in ffma[-3:]:
        add=re.findall(r'R(\d+)',t)[-1]
        # find defining FMUL
        for j in range(addr[a]-1,max(0,addr[a]-300),-1):
            tt=ins[j][1]
            mm=re.match(r'(@!?P\d+\s+)?FMUL\S*\s+R(\d+),',tt)
            if mm and mm.group(2)==add:
                if start is not None and int(ins[j][0],16)>=start: inside+=1
                else: outside+=1
                break
    res.append((ins[i][0],hex(start) if start else None,'B(y rounded)' if inside>outside else 'A(x rounded)',inside,outside))
for r in res: print(r)
