#!/usr/bin/env python
"""Spill stores inside the point loops of every k_frame instantiation of the built library (no GPU needed).

    python tools/spill_report.py [path/to/libgs_b200.so]

The projection loop sits at the 64-register cap of a 1024-thread CTA.  When ptxas spills a register of the PREFETCHED next
unit, the spill store right behind the load waits for that load and the prefetch is gone (found with ncu: one STL held 7 % of
all stall samples; removing it was worth +4.3 %).  Columns: the six instantiations of the loop per kernel (nearest / checked,
nearest / production, last / checked, last / production with tag shift 0, 1, any): spill stores / static instructions."""
import os, re, subprocess, sys, tempfile
lib=os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "groundsegmentationrosnode_b200", "libgs_b200.so")
tmp=tempfile.mkdtemp()
subprocess.run(["cuobjdump","-xelf","all",lib],cwd=tmp,check=True,stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL)
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    txt=subprocess.run(["nvdisasm","-g","-c",os.path.join(tmp,f)],capture_output=True,text=True).stdout
    fn=None; funcs={}; cur=""
    for ln in txt.splitlines():
        m=re.match(r"\s*\.text\.(\S+):",ln)
        if m: fn=m.group(1); continue
        if not(fn and "k_frame" in fn): continue
        m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
        if m: cur="%s:%s"%(os.path.basename(m.group(1)),m.group(2)); continue
        if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S",ln):
            funcs.setdefault(fn,[]).append((cur,re.sub(r"/\*[0-9a-f]{4,}\*/","",ln).strip()))
    for fn,rows in funcs.items():
        m=re.search(r"k_frameILi(\d)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELb(\d)",fn)
        name="<%s>"%",".join(m.groups()) if m else fn[:40]
        at=[i for i,(c,ins) in enumerate(rows) if ins.startswith("ATOMS.MAX")]
        res=[]
        for j in range(0,len(at)-1,2):
            a0,a1=at[j],at[j+1]
            ys=[i for i in range(max(0,a0-450),a0) if rows[i][1].startswith("YIELD")]
            if not ys: res.append("?"); continue
            st=ys[-1]-14
            en=a1
            while en < len(rows)-1 and not re.match(r"@!?U?P\d BRA",rows[en][1]): en+=1
            seg=rows[st:en+1]
            res.append("%d/%d" % (sum(1 for c,i in seg if "STL" in i), len(seg)))
        print(name, " spill stores / static size per point loop:", " ".join(res))
