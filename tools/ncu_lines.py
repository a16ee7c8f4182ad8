#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_frame<(int)1> [top_n]

ncu's `--page source --csv` lists SASS instructions in program order; nvdisasm -g gives the source line of every SASS
instruction of the same cubin in the same order.  Zipping the two attributes executed instructions / samples to lines of
our .cu files (inlined device functions are attributed to their own lines)."""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "groundsegmentationrosnode_b200", "libgs_b200.so")


def sass_lines(mangled_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line = None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if cur_fn and mangled_substr in cur_fn and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln):
                out.append(cur_line)
        if out:
            return out
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    mangled = sys.argv[4] if len(sys.argv) > 4 else "k_frameILi1"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + re.escape(kernel.split("<")[0])],
                         capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(raw))
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if len(r) != len(hdr):
            break
        body.append(dict(zip(hdr, r)))
    lines = sass_lines(mangled)
    if len(lines) != len(body):
        print("warning: %d SASS rows in the report, %d instructions in the disassembly" % (len(body), len(lines)))
    agg = collections.defaultdict(lambda: [0, 0])
    for d, l in zip(body, lines):
        agg[l][0] += int(d["Instructions Executed"] or 0)
        agg[l][1] += int(d["# Samples"] or 0)
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    src = {}
    print("total warp instructions %d, samples %d" % (ti, ts))
    for (l, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = ""
        if l:
            path = os.path.join(ROOT, "groundsegmentationrosnode_b200", "csrc", l[0])
            if os.path.exists(path):
                src.setdefault(path, open(path).read().splitlines())
                if l[1] - 1 < len(src[path]):
                    text = src[path][l[1] - 1].strip()[:100]
        print("%-22s inst %5.1f%%  samples %5.1f%%  %s" % ("%s:%d" % l if l else "?", 100.0 * v[0] / ti, 100.0 * v[1] / ts, text))


if __name__ == "__main__":
    main()
