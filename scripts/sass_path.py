#!/usr/bin/env python
"""Dynamic instruction count along ONE path through a kernel's SASS: walks a `cuobjdump -sass`
listing of a single function from a start address to an end address, following unconditional
branches and the conditional branches named as taken (addr or addr:times), and prints the
FP64 / other split -- 2 x FP64 + other is the issue-port time that bounds kernel 2 on B200.

    cuobjdump -sass perturb_b200/csrc/sgp4_prop.o | awk '/Function : .*kernelILi0ELi0ELb0ELi32E/{f=1} \
        f&&/Function : /&&!/ILi0ELi0ELb0ELi32E/{f=0} f' > /tmp/k0.sass
    python scripts/sass_path.py /tmp/k0.sass a80 3c40 1020,2ff0,3230,3780
(profiles/r2_sass_mix.txt holds the result for the shipped near-Earth kernel.)"""
import re,sys,collections
f=sys.argv[1]; start=int(sys.argv[2],16); end=int(sys.argv[3],16)
taken={}
for x in (sys.argv[4].split(',') if len(sys.argv)>4 and sys.argv[4] else []):
    a,_,n=x.partition(':'); taken[int(a,16)]=int(n) if n else 10**9
ins={}
for line in open(f):
    m=re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);",line)
    if m: ins[int(m.group(1),16)]=m.group(2).strip()
FP64=("DFMA","DMUL","DADD","DSETP")
pc=start; c=collections.Counter(); n=0
while True:
    t=ins[pc]; ops=t.split(); op=ops[1] if ops[0].startswith("@") else ops[0]
    c[op.split('.')[0]]+=1; n+=1
    if pc==end: break
    if op.startswith("BRA"):
        tgt=int(re.search(r"0x([0-9a-f]+)",t).group(1),16)
        uncond=not ops[0].startswith("@") and "UP" not in t
        if uncond: pc=tgt; continue
        if taken.get(pc,0)>0: taken[pc]-=1; pc=tgt; continue
    pc+=16
f64=sum(c[k] for k in FP64)
print("%d instr: %d FP64 + %d other -> 2F+O = %d"%(n,f64,n-f64,2*f64+n-f64))
print(", ".join("%s %d"%kv for kv in c.most_common(30)))
