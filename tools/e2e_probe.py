#!/usr/bin/env python
"""Where the end-to-end (host buffers in, host buffers out) rate of 1..N GPUs comes from: the box's host-copy ceiling next to
the native multi-GPU driver (gs_multi: one host thread + context per GPU, NUMA-local pinned buffers).  One process.

    gpurun --gpus 8 -- 'python tools/e2e_probe.py > gpurun_out/e2e_probe.txt'

For G in 1, 2, 4, 8 (as many GPUs as are visible):
  raw      concurrent pinned-host <-> device copies on G GPUs: H2D alone, D2H alone, both directions at once
           (a) buffers allocated by the main thread (wherever the OS puts them), (b) buffers first-touched by the worker bound next to each GPU
  e2e      gs_multi_process_batch_host on G GPUs: frames/s, and the H2D rate that implies, with 16-byte XYZI and 12-byte XYZ records
"""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from groundsegmentationrosnode_b200 import capi, params as P, synth  # noqa: E402


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:
        return "(%s)" % e


def raw_copies(G, host_bufs_in, host_bufs_out, n_bytes, reps=4):
    """host_bufs_*: per GPU a pinned uint8 torch tensor of n_bytes"""
    dev = [torch.empty(n_bytes, dtype=torch.uint8, device="cuda:%d" % g) for g in range(G)]
    dev2 = [torch.empty(n_bytes, dtype=torch.uint8, device="cuda:%d" % g) for g in range(G)]
    s_in = [torch.cuda.Stream(device=g) for g in range(G)]
    s_out = [torch.cuda.Stream(device=g) for g in range(G)]
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for g in range(G):
            torch.cuda.synchronize(g)
        t0 = time.perf_counter()
        for _ in range(reps):
            for g in range(G):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s_in[g]):
                        dev[g].copy_(host_bufs_in[g], non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s_out[g]):
                        host_bufs_out[g].copy_(dev2[g], non_blocking=True)
        for g in range(G):
            torch.cuda.synchronize(g)
        dt = time.perf_counter() - t0
        res[mode] = reps * n_bytes * G / dt / 1e9  # per direction, summed over GPUs
    return res


def main():
    n_dev = torch.cuda.device_count()
    print("GPUs visible: %d, host CPUs: %d (affinity %d)" % (n_dev, os.cpu_count(), len(os.sched_getaffinity(0))))
    print("NUMA nodes:", sh("ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l"), "|", sh("lscpu | grep -i 'numa\\|model name\\|socket' | tr '\\n' ';'"))
    print(sh("nvidia-smi topo -m 2>/dev/null | head -14"))
    p = P.preset(P.GS_PRESET_HDL64E_2048)
    H, W = p.shape
    pool, offs = synth.make_pool(p, 16)
    n_raw = 256 << 20
    for G in [g for g in (1, 2, 4, 8) if g <= n_dev]:
        F = 512 * G
        pts, offsets = synth.tile_pool(pool, offs, F)
        with capi.MultiGpu(p, list(range(G))) as mg:
            print("\n=== %d GPU(s): workers %s" % (G, mg.workers()))
            # raw copy ceilings
            plain_in = [torch.empty(n_raw, dtype=torch.uint8).pin_memory() for _ in range(G)]
            plain_out = [torch.empty(n_raw, dtype=torch.uint8).pin_memory() for _ in range(G)]
            r = raw_copies(G, plain_in, plain_out, n_raw)
            print("raw copies, buffers pinned by the main thread : H2D %.1f GB/s, D2H %.1f GB/s, both at once %.1f GB/s per direction (sum over GPUs)" % (r["h2d"], r["d2h"], r["both"]))
            del plain_in, plain_out
            big_in = mg.host_array((G * n_raw,), np.uint8)   # block g first-touched by worker g
            big_out = mg.host_array((G * n_raw,), np.uint8)
            loc_in = [torch.from_numpy(big_in[g * n_raw:(g + 1) * n_raw]) for g in range(G)]
            loc_out = [torch.from_numpy(big_out[g * n_raw:(g + 1) * n_raw]) for g in range(G)]
            r = raw_copies(G, loc_in, loc_out, n_raw)
            print("raw copies, buffers first-touched next to GPU : H2D %.1f GB/s, D2H %.1f GB/s, both at once %.1f GB/s per direction (sum over GPUs)" % (r["h2d"], r["d2h"], r["both"]))
            raw_both = r["both"]
            # end to end through the native multi-GPU driver
            h_res = mg.host_array((F, H, W), np.uint16)
            h_xyz = mg.host_array((F * 70000, 4), np.float32)
            for name, arr in (("XYZI 16 B", pts), ("XYZ 12 B", np.ascontiguousarray(pts[:, :3]))):
                h_pts = mg.host_array(arr.shape, np.float32)
                h_pts[:] = arr
                mg.process_batch_host(h_pts, offsets, h_res, h_xyz)
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    _, _, first, count = mg.process_batch_host(h_pts, offsets, h_res, h_xyz)
                dt = (time.perf_counter() - t0) / reps
                in_bytes = arr.nbytes
                out_bytes = h_res.nbytes + int(count.sum()) * 16
                print("e2e %-10s: %8.0f frames/s on %d GPU(s) (%6.0f per GPU)  H2D %.1f GB/s + D2H %.1f GB/s  = %.0f %% of the raw both-directions H2D rate"
                      % (name, F / dt, G, F / dt / G, in_bytes / dt / 1e9, out_bytes / dt / 1e9, 100.0 * (in_bytes / dt / 1e9) / raw_both))
            # input only / no cloud: which direction binds
            t0 = time.perf_counter()
            mg.process_batch_host(h_pts, offsets, h_res, None)
            dt = time.perf_counter() - t0
            print("e2e XYZ 12 B, image only (no cloud out): %8.0f frames/s" % (F / dt))


if __name__ == "__main__":
    main()
