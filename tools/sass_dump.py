#!/usr/bin/env python
"""SASS of one kernel of libgs_b200.so, every instruction tagged with its source line (and inlining chain) — no GPU needed.

    python tools/sass_dump.py [kernel-substring] [out.txt]      default: the C2 / C4 instantiation k_frame<1,1024,64,512,4>

Also prints the instruction count per phase of the fused kernel (source ranges between the GS_PHASE marks) and per
inlined helper, which is what the per-frame instruction budget is read from.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("GS_LIB", os.path.join(ROOT, "groundsegmentationrosnode_b200", "libgs_b200.so"))
sub = sys.argv[1] if len(sys.argv) > 1 else "k_frameILi1ELi1024ELi64ELi512ELi4"
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "sass_%s.txt" % re.sub(r"\W", "_", sub))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
rows = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    fn, cur = None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
            continue
        if fn and sub in fn and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln):
            ins = re.sub(r"/\*[0-9a-f]{4,}\*/", "", ln).strip()
            rows.append((cur, ins))
os.makedirs(os.path.dirname(out), exist_ok=True)
with open(out, "w") as fh:
    for i, (c, ins) in enumerate(rows):
        fh.write("%6d %-28s %s   %s\n" % (i, "%s:%d" % (c[0], c[1]) if c else "?", ins, c[2] if c else ""))
print("%d instructions -> %s" % (len(rows), out))
# static mix
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0] for _, ins in rows)
print("opcode mix (static):", ", ".join("%s %d" % kv for kv in ops.most_common(25)))
