#!/usr/bin/env python
"""Device-resident throughput of the fused path for the BASELINE sensor shapes (needs a GPU).

    python tools/preset_throughput.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from groundsegmentationrosnode_b200 import capi, params as P, synth  # noqa: E402

for name, preset, F in (("C1 VLP-16 16x1800", P.GS_PRESET_VLP16, 4096), ("C2 HDL-64E 64x2048", P.GS_PRESET_HDL64E_2048, 2048),
                        ("C3 OS1-128 128x2048", P.GS_PRESET_OS1_128, 1024), ("cloud2range default 64x2000 (SG-7)", P.GS_PRESET_CLOUD2RANGE_DEFAULT, 2048),
                        ("Alfa_GS default 64x1800", P.GS_PRESET_ALFA_GS_DEFAULT, 2048)):
    p = P.preset(preset)
    pool, offs = synth.make_pool(p, 16)
    pts, offsets = synth.tile_pool(pool, offs, F)
    H, W = p.shape
    with capi.Context(p) as ctx:
        d_pts = torch.from_numpy(pts).cuda()
        d_res = torch.empty((F, H, W), dtype=torch.int16, device="cuda")
        d_xyz = torch.empty((F, H * W, 4), dtype=torch.float32, device="cuda")
        d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        for _ in range(3):
            ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
        ctx.synchronize()
        l0 = ctx.stats().kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(5):
                ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
            e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        n_ng = int(d_cnt.sum().item())
        nbytes = 16.0 * int(offsets[-1]) + 2.0 * H * W * F + 16.0 * n_ng
        print("%-36s %9.0f frames/s  %.3e points/s  %6.1f GB/s algorithmic  (%d launches per batch of %d frames)" % (
            name, F / ms * 1e3, int(offsets[-1]) / ms * 1e3, nbytes / ms / 1e6, (ctx.stats().kernel_launches - l0) // 5, F), flush=True)
        del d_pts, d_res, d_xyz, d_cnt
        torch.cuda.empty_cache()
