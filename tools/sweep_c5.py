#!/usr/bin/env python
"""BASELINE config C5: angle-threshold sweep on 128 x 2048 frames (needs a GPU).

Ground threshold 2..10 degrees: device-resident throughput of the whole path (fused kernel) and mask parity with the oracle
on one frame; Labeler threshold beta 5..20 degrees: the whole path plus the all-components CCL on every frame of the batch
(gs_process_batch_device_labels), throughput and label parity with the oracle.

    python tools/sweep_c5.py
"""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from groundsegmentationrosnode_b200 import capi, params as P, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)

base = P.preset(P.GS_PRESET_OS1_128)
H, W = base.shape
F = 480
pool, offs = synth.make_pool(base, 8)
pts, offsets = synth.tile_pool(pool, offs, F)
d_pts = torch.from_numpy(pts).cuda()
d_res = torch.empty((F, H, W), dtype=torch.int16, device="cuda")
d_xyz = torch.empty((F, H * W, 4), dtype=torch.float32, device="cuda")
d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
frame0 = pool[: offs[1]]
print("ground threshold sweep, %d frames of %dx%d per batch" % (F, H, W))
for deg in (2, 3, 4, 5, 6, 7, 8, 9, 10):
    p = base.copy(ground_angle_threshold=math.radians(deg))
    want = O.process_frame(p, frame0)
    with capi.Context(p) as ctx:
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        for _ in range(2):
            ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(3):
                ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
            e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        same = np.array_equal(d_res[0].cpu().numpy().view(np.uint16), want["no_ground"]) and int(d_cnt[0]) == want["xyz"].shape[0]
        print("  ground %2d deg (cut-off K = %2d): %8.0f frames/s, %.3e points/s, mask parity with the oracle: %s" % (
            deg, capi.derive(p).k_ground, F / ms * 1e3, int(offsets[-1]) / ms * 1e3, "exact" if same else "MISMATCH"), flush=True)
print("Labeler beta sweep: the whole path plus the all-components CCL on every frame (gs_process_batch_device_labels), %d frames per batch" % F)
d_lab = torch.empty((F, H, W), dtype=torch.int32, device="cuda")
d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
n_distinct = len(offs) - 1
o = [O.process_frame(base, pool[offs[i]:offs[i + 1]]) for i in range(2)]
with capi.Context(base) as ctx:
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    for deg in (5, 10, 15, 20):
        beta = math.radians(deg)
        want = [O.label_components(base, x["no_ground"], x["angle"], beta) for x in o]

        def step():
            ctx.process_batch_device_labels(d_pts.data_ptr(), offsets, F, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr(), d_lab.data_ptr(), beta,
                                            d_n_components=d_n.data_ptr())
        for _ in range(2):
            step()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(3):
                step()
            e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        same = all(np.array_equal(d_lab[i].cpu().numpy(), want[i][0]) and int(d_n[i]) == want[i][1] and
                   np.array_equal(d_lab[i + n_distinct].cpu().numpy(), want[i][0]) for i in range(2))
        print("  beta %2d deg (cut-off K = %2d): %6d components per frame, %8.0f frames/s, %.3e points/s, labels equal to the oracle's canonical labels: %s" % (
            deg, capi.derive(base.copy(ground_angle_threshold=beta)).k_ground, want[0][1], F / ms * 1e3, int(offsets[-1]) / ms * 1e3, "yes" if same else "NO"), flush=True)
