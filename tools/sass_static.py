#!/usr/bin/env python
"""Static SASS listing of one kernel of libgs_b200.so with source lines (no GPU, no profile needed).

    python tools/sass_static.py k_frameILi1ELi1024ELi64 [first_line last_line]
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "groundsegmentationrosnode_b200", "libgs_b200.so")
sub = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur = None, None
    n = 0
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if cur_fn and sub in cur_fn and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln):
            n += 1
            if cur and cur[0] == "gs_fused.cu" and lo <= cur[1] <= hi:
                ins = re.sub(r"/\*[0-9a-f]{4,}\*/", "", ln).strip()
                print("%5d %s:%d  %s" % (n, cur[0], cur[1], ins[:100]))
