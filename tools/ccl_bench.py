#!/usr/bin/env python
"""The batched Labeler CCL (gs_label_components_batch_device) timed on its own, next to the fused pipeline that feeds it (needs a GPU).

    python tools/ccl_bench.py [preset] [frames]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from groundsegmentationrosnode_b200 import capi, params as P, synth  # noqa: E402

preset = int(sys.argv[1]) if len(sys.argv) > 1 else P.GS_PRESET_OS1_128
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
p = P.preset(preset)
H, W = p.shape
pool, offs = synth.make_pool(p, 8)
pts, offsets = synth.tile_pool(pool, offs, F)
beta = float(np.radians(10.0))
with capi.Context(p) as ctx:
    d_pts = torch.from_numpy(pts).cuda()
    d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
    d_ang = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
    d_lab = torch.zeros((F, H, W), dtype=torch.int32, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    def timed(fn, reps=5):
        fn()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
        ctx.synchronize()
        return e0.elapsed_time(e1) / reps

    t_pipe = timed(lambda: ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, d_ang.data_ptr(), d_res.data_ptr(), 0, 0))
    t_ccl = timed(lambda: ctx.label_components_batch_device(d_res.data_ptr(), d_ang.data_ptr(), F, beta, d_lab.data_ptr(), d_n.data_ptr()))
    t_both = timed(lambda: ctx.process_batch_device_labels(d_pts.data_ptr(), offsets, F, d_res.data_ptr(), 0, 0, d_lab.data_ptr(), beta))
    print("%dx%d, %d frames: pipeline (image + angle out) %.2f us/frame, CCL alone %.2f us/frame (%d components/frame), both in waves %.2f us/frame = %.0f frames/s"
          % (H, W, F, 1e3 * t_pipe / F, 1e3 * t_ccl / F, int(d_n.float().mean().item()), 1e3 * t_both / F, F / t_both * 1e3))
