#!/usr/bin/env python
"""sass_sim.py -- developer tool: follow the hot path of a kernel through its SASS and estimate how it issues on an SMSP.

    cuobjdump -sass -fun <mangled kernel> kt_kernels_exact.o > k.sass
    python scripts/sass_sim.py k.sass auto | 0x340 [0x420=T ...]   # start address (the main-loop head), branch hints addr=T|N

trace(): follows instructions from the start address; forward conditional branches are not taken unless they skip a short
block that ends in a CALL or a far jump (cold blocks: slow-path subroutines, early exits), a hint says otherwise, and the walk
stops at the first taken branch back to (or before) the start address -- i.e. after one iteration of the main loop.
simulate(): N warps run that instruction list in a loop on one SMSP: in-order issue per warp, one issue per cycle, the FP64 pipe
busy 2 cycles per warp instruction, fixed result latencies (FP64 8, ALU 4.5, MUFU 18, conversions 14, local loads 28), a
12-cycle bubble after a branch, 6 after BSSY / BSYNC.  Calibration on B200 (round 1): the model gives 5 350 cycles per warp-step
for the exact RK4 kernel at 4 warps per SMSP where ncu's per-instruction samples give 6 270; it ranks variants correctly when they
differ in instruction count or branch structure and is blind to register-bank and dispatch effects.
Also prints the opcode histogram of the path and the latency-weighted dataflow critical path (the in-order schedule's lower bound).
"""
import collections
import re
import sys

def parse(f):
    ins=[]
    for l in open(f):
        m=re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
        if m: ins.append((int(m.group(1),16),m.group(2).strip()))
    return ins
LAT={'D':8.0,'MUFU':18,'I2F':14,'F2I':14,'LDL':28,'LDS':24,'LDC':12,'LDCU':12,'LDG':300,'S2R':20,'SHFL':24,'ATOM':40,'POPC':10,'FLO':10}
def klass(op):
    o=op.split('.')[0]
    if o in('DFMA','DMUL','DADD','DSETP'): return 'D'
    return o
def regs_of(tok,wide):
    out=[]
    for m in re.finditer(r'(?<![A-Za-z0-9_])(U?R|U?P)(\d+)',tok):
        k,n=m.group(1),int(m.group(2))
        out.append((k,n))
        if wide and k in('R','UR'): out.append((k,n+1))
    return out
def decode(t):
    pred=None
    m=re.match(r'@(!?)(U?P\d+)\s+(.*)',t)
    if m: pred=m.group(2); t=m.group(3)
    parts=t.split(None,1)
    op=parts[0]; rest=parts[1] if len(parts)>1 else ''
    ops=[x.strip() for x in rest.split(',')] if rest else []
    k=klass(op)
    wide = k=='D' or '.64' in op or 'WIDE' in op or op.startswith('I2F.F64') 
    srcwide = k=='D' or op.startswith('F2I.F64')
    dst=[];src=[]
    nodst = op.startswith(('ST','BRA','BSSY','BSYNC','EXIT','BREAK','YIELD','CALL','RET','NOP','WARPSYNC','BAR','RED','MEMBAR'))
    for i,o in enumerate(ops):
        if i==0 and not nodst:
            w = wide and not op.startswith(('DSETP','MUFU'))
            if op.startswith(('DSETP','ISETP','FSETP','PLOP3')):
                dst+=regs_of(o,False)
                continue
            dst+=regs_of(o,w)
        elif i==1 and op.startswith(('DSETP','ISETP','FSETP','PLOP3')):
            dst+=regs_of(o,False)
        else:
            src+=regs_of(o,srcwide and not o.startswith(('P','!P','UP')) )
    if op.startswith('MUFU.RCP64H') or op.startswith('MUFU.RSQ64H'):
        pass
    if pred: src.append((pred[0:-len(re.search(r'\d+$',pred).group())],int(re.search(r'\d+$',pred).group())))
    return op,k,dst,src
def linear(ins,lo,hi,skips):
    out=[]
    for a,t in ins:
        if a<lo or a>hi: continue
        if any(s<=a<=e for s,e in skips): continue
        out.append((a,t))
    return out
def simulate(prog,nwarps,iters=6,verbose=False,brpen=8,syncpen=4,dlat=8.0):
    dec=[decode(t) for a,t in prog]
    n=len(dec)
    nfp=sum(1 for d in dec if d[1]=='D')
    # state per warp
    pc=[0]*nwarps; it=[0]*nwarps
    ready=[dict() for _ in range(nwarps)]
    nextissue=[w*7 for w in range(nwarps)]
    fp_free=0; cyc=0; done=0
    start_c=None; fin=[0]*nwarps
    last=0
    total=n*iters*nwarps
    issued=0
    rr=0
    stall_hist={}
    while issued<total:
        # pick a warp that can issue at this cycle
        picked=False
        for j in range(nwarps):
            w=(rr+j)%nwarps
            if it[w]>=iters: continue
            if nextissue[w]>cyc: continue
            op,k,dst,src=dec[pc[w]]
            t=max([ready[w].get(r,0) for r in src]+[0])
            if t>cyc: continue
            if k=='D' and fp_free>cyc: continue
            # issue
            lat=dlat if k=='D' else LAT.get(k,4.5)
            for r in dst: ready[w][r]=cyc+lat
            if k=='D': fp_free=cyc+2
            nextissue[w]=cyc+1
            if k=='BRA': nextissue[w]=cyc+brpen
            elif k in('BSYNC','BSSY'): nextissue[w]=cyc+syncpen
            pc[w]+=1
            if pc[w]==n:
                pc[w]=0; it[w]+=1; fin[w]=cyc
            issued+=1; picked=True; rr=(w+1)%nwarps
            break
        cyc+=1
    return cyc, n, nfp
def build(ins):
    idx={a:i for i,(a,t) in enumerate(ins)}
    return idx
def has_call(ins,idx,a,b):
    i=idx[a]+1
    while i<len(ins) and ins[i][0]<b:
        if 'CALL' in ins[i][1]: return True
        i+=1
    return False
def has_far(ins,idx,a,b):
    i=idx[a]+1
    while i<len(ins) and ins[i][0]<b:
        t=ins[i][1]
        if 'BRA' in t:
            m=re.search(r'0x([0-9a-f]+)',t)
            if m and (int(m.group(1),16)-ins[i][0])//16>600: return True
        i+=1
    return False
def trace(ins,idx,start,stops,hints,maxn=5000,log=False):
    """hints: dict addr->bool (taken) ; or predicate-name -> bool via 'pred:P5'"""
    out=[];a=start
    while len(out)<maxn:
        i=idx[a]; t=ins[i][1]
        out.append((a,t))
        nxt=ins[i+1][0] if i+1<len(ins) else None
        if re.search(r'\bBRA\b',t.split()[0] if not t.startswith('@') else t.split()[1]) or ' BRA' in ' '+t:
            m=re.search(r'0x([0-9a-f]+)',t)
            tgt=int(m.group(1),16) if m else None
            pm=re.match(r'@(!?)(U?P\d+)\s+BRA(?:\.U)?\s+(?:(!?)(U?P\d+),\s*)?',t)
            um=re.match(r'BRA\.U\s+(!?)(UP\d+)',t)
            if t.startswith('BRA') and not um:   # unconditional
                taken=True
            else:
                if a in hints: taken=hints[a]
                else:
                    key=None
                    if pm: neg,p=pm.group(1),pm.group(2)
                    else: neg,p=um.group(1),um.group(2)
                    if ('pred:'+p) in hints and not (pm and pm.group(4)):
                        v=hints['pred:'+p]; taken = (not v) if neg else v
                    elif tgt>a and (tgt-a)//16<70 and has_call(ins,idx,a,tgt): taken=True
                    elif tgt>a and (tgt-a)//16<70 and has_far(ins,idx,a,tgt): taken=True
                    elif tgt<=a: taken=False
                    else: taken=False
                if log: print(hex(a),t,'->','T' if taken else 'N')
            if taken:
                if tgt in stops or tgt<=start: break
                a=tgt; continue
        if nxt in stops: break
        if 'EXIT' in t or 'RET' in t: break
        a=nxt
    return out


if __name__ == "__main__":
    f = sys.argv[1]
    ins = parse(f)
    idx = build(ins)
    hints = {}
    if sys.argv[2] == "auto":
        # trace_persistent (round 2): the stepping loop is the innermost loop that holds a whole integration step, i.e. the
        # backward branch with the smallest span above 300 instructions; its first conditional branch skips the step for a lane
        # without an active ray (not taken in the steady state)
        back = []
        for a, t in ins:
            m = re.search(r"0x([0-9a-f]+)", t)
            if "BRA" in t and m and int(m.group(1), 16) < a and (a - int(m.group(1), 16)) // 16 > 300:
                tgt = int(m.group(1), 16)
                if any("VOTE" in u for b, u in ins if tgt <= b <= a):     # (the loop's exit vote: rules out loops of cold subroutines)
                    back.append(((a - tgt) // 16, tgt))
        start = min(back)[1]
        print(f"stepping-loop head {start:#x} ({min(back)[0]} instructions to its backward branch)")
    else:
        start = int(sys.argv[2], 16)
    for h in sys.argv[3:]:
        k, v = h.split("=")
        hints[int(k, 16)] = (v == "T")
    prog = trace(ins, idx, start, set(), hints, maxn=4000)
    c = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for a, t in prog)
    print(f"{len(prog)} instructions on the path; FP64 {c['DFMA'] + c['DMUL'] + c['DADD']} (+{c['DSETP']} DSETP), MUFU {c['MUFU']}, "
          f"BRA {c['BRA']}, BSSY {c['BSSY']}, LDL {c['LDL']}, STL {c['STL']}")
    print(sorted(c.items(), key=lambda x: -x[1])[:16])
    for nw in (1, 2, 3, 4, 5, 6, 8):
        cyc, n, nfp = simulate(prog, nw, 5, brpen=12, syncpen=6, dlat=8)
        print(f"{nw} warps/SMSP: {cyc / 5:8.0f} cycles per pass per warp, {cyc / 5 / nw:8.0f} per pass per SMSP "
              f"(FP64 pipe bound {2 * (nfp):d})")
    ready, finish = {}, 0
    for op, k, dst, src in (decode(t) for a, t in prog):
        t0 = max([ready.get(r, 0) for r in src] + [0])
        lat = 8.0 if k == "D" else LAT.get(k, 4.5)
        for r in dst:
            ready[r] = t0 + lat
        finish = max(finish, t0 + lat)
    print(f"dataflow critical path: {finish:.0f} cycles")
