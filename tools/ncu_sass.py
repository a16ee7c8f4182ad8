#!/usr/bin/env python
"""SASS listing of one kernel from an .ncu-rep with executed counts and source lines (no GPU needed).

    python tools/ncu_sass.py gpurun_out/prof.ncu-rep k_frameILi1ELi1024ELi64 > /tmp/sass.txt
"""
import csv
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as T  # noqa: E402

rep, mangled = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:k_frame"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
body = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr)]
lines = T.sass_lines(mangled)
for i, (d, l) in enumerate(zip(body, lines)):
    print("%5d %-22s %10s %6s  %s" % (i, "%s:%d" % l if l else "?", d["Instructions Executed"], d["# Samples"], d["Source"]))
