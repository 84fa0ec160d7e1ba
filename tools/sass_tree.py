"""Expression trees out of SASS (test tooling for the parity work; reads cuobjdump -sass text, no GPU needed).

For a texture fetch of a kernel, follows the registers of its coordinates backwards through FMUL / FFMA / FADD /
MUFU / I2FP / MOV to the loads they come from and prints the arithmetic as a nested expression, commutative
operands sorted.  Two call sites compute the same value bit for bit iff their trees are equal; this is how the
reference's inlined copies of ComputeBilateralNCCOld were compared with each other and with this repository's
restatement (5 of its 12 copies in the sweep kernel round the x products of a tap coordinate,
7 the y products).

  cuobjdump -sass -fun <kernel> oracle/_ref/ref_stage_probe > k.sass
  python tools/sass_tree.py k.sass <address of the TEX instruction, e.g. 10b0> [depth]
"""
import re,sys
sys.setrecursionlimit(100000)
def load(path):
    ins=[]
    for l in open(path):
        m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l)
        if m: ins.append((m.group(1),m.group(2).strip()))
    return ins
def parse(t):
    pred=None
    m=re.match(r'^(@!?U?P\d+)\s+(.*)$',t)
    if m: pred,t=m.group(1),m.group(2)
    op,rest=(t.split(None,1)+[''])[:2]
    ops=[o.strip() for o in rest.split(',')] if rest else []
    return pred,op,ops
ARITH={'FMUL','FFMA','FADD','MUFU','I2FP','MOV','FSEL'}
def src_tok(o):
    m=re.match(r'^(-?)(\|?)R(\d+)(\.reuse)?(\|?)$',o)
    if m: return ('reg',m.group(3),m.group(1)=='-',m.group(2)=='|')
    return ('imm',o,False,False)
def build(ins,idx,reg,depth,memo):
    """expression tree (as a string) of the value of `reg` just before instruction idx"""
    key=(idx,reg)
    if key in memo: return memo[key]
    res='R?'
    for j in range(idx-1,max(-1,idx-4000),-1):
        pred,op,ops=parse(ins[j][1])
        if not ops: continue
        base=op.split('.')[0]
        dst=re.match(r'^R(\d+)$',ops[0])
        if base in ('LDG','LD','LDL','LDS','LDC','LDCU'):
            # 64/128-bit loads define consecutive regs
            w=2 if '.64' in op else (4 if '.128' in op else 1)
            if dst and int(dst.group(1))<=int(reg)<int(dst.group(1))+w:
                off=int(reg)-int(dst.group(1))
                addr=ops[1] if len(ops)>1 else '?'
                m=re.search(r'\+(-?0x[0-9a-f]+)\]',addr)
                o=int(m.group(1),16) if m else 0
                res=f'{base}[{hex(o+4*off)}]'
                break
            continue
        if not dst or dst.group(1)!=reg: continue
        if base not in ARITH or depth==0:
            res=f'{base}@{ins[j][0]}'; break
        args=[]
        for o in ops[1:]:
            k=src_tok(o)
            if k[0]=='reg':
                sub=build(ins,j,k[1],depth-1,memo)
                if k[3]: sub='|'+sub+'|'
                if k[2]: sub='-'+sub
                args.append(sub)
            else: args.append(k[1])
        name=op.replace('.FTZ','').replace('.F32.S32','')
        if base in ('FMUL','FADD'): args=sorted(args)
        if base=='FFMA': args=sorted(args[:2])+args[2:]
        res=f'{name}({",".join(args)})'
        break
    memo[key]=res
    return res
if __name__=='__main__':
    path,texaddr=sys.argv[1],sys.argv[2]
    depth=int(sys.argv[3]) if len(sys.argv)>3 else 40
    ins=load(path)
    ti=[i for i,(a,t) in enumerate(ins) if a==texaddr][0]
    _,op,ops=parse(ins[ti][1])
    coord=re.findall(r'R(\d+)',ops[2])[0]
    memo={}
    print('U =',build(ins,ti,coord,depth,memo))
    print('V =',build(ins,ti,str(int(coord)+1),depth,memo))
