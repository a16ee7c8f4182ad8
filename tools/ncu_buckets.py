#!/usr/bin/env python
"""Instruction / stall-sample shares of the fused kernel by phase (source line ranges are read from the GS_PHASE marks).

    python tools/ncu_buckets.py gpurun_out/prof.ncu-rep [frames]
"""
import collections
import csv
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as T  # noqa: E402

rep = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 296
src = open(os.path.join(T.ROOT, "groundsegmentationrosnode_b200", "csrc", "gs_fused.cu")).read().splitlines()
marks = [(i + 1, int(re.search(r"GS_PHASE\((\d+)\)", l).group(1))) for i, l in enumerate(src) if re.match(r"\s+GS_PHASE\(\d+\);", l)]
kernel_start = [i + 1 for i, l in enumerate(src) if "__global__ void __launch_bounds__(NT" in l][0]
names = {0: "clear+deal", 1: "project rounds", 2: "wait rounds", 3: "remote+compact", 4: "angle walk", 5: "wait angles", 6: "edges+seeds", 7: "wait edges",
         8: "link scan", 9: "boundary rounds", 10: "readout", 11: "label/dilate/transpose", 12: "wait rowcounts", 13: "output", 14: "unions", 15: "seed flags"}


def bucket(l):
    if l is None:
        return "other"
    f, n = l
    if f != "gs_fused.cu":
        return "lib:" + f
    if n < kernel_start:
        for i in range(n - 1, -1, -1):
            m = re.match(r"__device__ \S+ (?:\S+ )?(\w+)\(", src[i]) or re.match(r"(?:template.*)?struct (\w+)", src[i])
            if m:
                return "fn:" + m.group(1)
        return "helpers"
    cur = "prologue"
    if n >= marks[0][0]:
        cur = "clear+deal"
    for ln, ph in marks:
        if n >= ln:
            cur = names.get(ph, str(ph))
    return cur


raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:k_frame"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:  # the first kernel of the report only (the cluster instantiation; its software-cluster twin follows)
    if len(r) != len(hdr) or r[0] == "Address" or r[0] == "Kernel Name":
        break
    body.append(dict(zip(hdr, r)))
lines = T.sass_lines(sys.argv[3] if len(sys.argv) > 3 else "k_frameILi1ELi1024ELi64ELi512ELi4ELb0")
STALLS = ["stall_barrier", "stall_short_sb", "stall_wait", "stall_long_sb", "stall_not_selected", "stall_branch_resolving", "stall_math", "stall_mio"]
if len(lines) != len(body):
    print("warning: %d SASS rows in the report, %d instructions in the library on disk: the capture is of another build, the attribution below is off" % (len(body), len(lines)))
agg = collections.defaultdict(lambda: [0, 0] + [0] * len(STALLS))
thr = collections.defaultdict(int)
for d, l in zip(body, lines):
    b = bucket(l)
    agg[b][0] += int(d["Instructions Executed"] or 0)
    agg[b][1] += int(d["# Samples"] or 0)
    thr[b] += int(d.get("Thread Instructions Executed") or 0)
    for k, st in enumerate(STALLS):
        agg[b][2 + k] += int(d.get(st) or 0)
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print("total %.2f M warp-instructions per frame" % (ti / frames / 1e6))
print("%-28s %-26s %-9s | share of all samples by stall reason: %s" % ("phase", "instructions", "samples", " ".join(x.replace("stall_", "")[:8] for x in STALLS)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] * 200 < ts:
        continue
    print("%-28s inst %5.1f%% (%7.1fk/frame, %4.1f lanes)   samples %5.1f%% | %s" % (k, 100 * v[0] / ti, v[0] / frames / 1e3, thr[k] / max(v[0], 1),
                                                                                  100 * v[1] / ts, " ".join("%8.1f" % (100.0 * x / ts) for x in v[2:])))
