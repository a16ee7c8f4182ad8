#!/usr/bin/env python
"""Throughput of the paths beside the headline one, 64x2048 frames, device-resident batches (needs a GPU):
  ordered (firing order, the headline)  |  row-major organised clouds  |  shuffled clouds (every point through the remote list)
  pixel_policy = nearest (fused kernel)  |  the per-stage kernels on the default configuration

    python tools/slow_paths.py [frames]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from groundsegmentationrosnode_b200 import capi, params as P, synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 592
base = P.preset(P.GS_PRESET_HDL64E_2048)
H, W = base.shape
pool, offs = synth.make_pool(base, 8)


def reorder(kind):
    frames = []
    rng = np.random.default_rng(3)
    for i in range(len(offs) - 1):
        f = pool[offs[i]:offs[i + 1]]
        if kind == "shuffled":
            f = f[rng.permutation(f.shape[0])]
        elif kind == "row-major":
            # organised cloud: beam by beam (row-major over the range image) instead of firing order (column by column)
            alt = np.arctan2(f[:, 2], np.hypot(f[:, 0], f[:, 1]))
            az = np.mod(np.arctan2(f[:, 1], f[:, 0]), 2 * np.pi)
            row = np.round((base.max_angle - alt) / (base.max_angle - base.min_angle) * (H - 1)).astype(np.int64)
            f = f[np.lexsort((az, row))]
        frames.append(np.ascontiguousarray(f))
    o = np.zeros(len(frames) + 1, np.int64)
    o[1:] = np.cumsum([x.shape[0] for x in frames])
    return np.concatenate(frames), o


def run(name, p, pts_pool, o_pool, staged=False):
    pts, offsets = synth.tile_pool(pts_pool, o_pool, F)
    with capi.Context(p) as ctx:
        ctx.set_option(capi.GS_OPT_FORCE_STAGED, int(staged))
        d_pts = torch.from_numpy(pts).cuda()
        d_res = torch.empty((F, H, W), dtype=torch.int16, device="cuda")
        d_xyz = torch.empty((F, H * W, 4), dtype=torch.float32, device="cuda")
        d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        for _ in range(2):
            ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
        ctx.synchronize()
        ctx.set_option(capi.GS_OPT_COLLECT_STATS, 0)
        l0 = ctx.stats().kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(3):
                ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
            e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("%-58s %9.0f frames/s  (%d launches per batch of %d frames)" % (name, F / ms * 1e3, (ctx.stats().kernel_launches - l0) // 3, F), flush=True)


run("ordered cloud (firing order), fused kernel", base, pool, offs)
rm, rmo = reorder("row-major")
run("row-major organised cloud, fused kernel", base, rm, rmo)
sh, sho = reorder("shuffled")
run("shuffled cloud (remote-list path), fused kernel", base, sh, sho)
run("pixel_policy = nearest (north-star variant), fused kernel", base.copy(pixel_policy=P.GS_PIXEL_NEAREST), pool, offs)
run("default configuration forced onto the per-stage kernels", base, pool, offs, staged=True)
