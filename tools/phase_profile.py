#!/usr/bin/env python
"""Per-phase SM-cycle breakdown of the fused kernel on a batch of synthetic frames (needs a GPU).

    python tools/phase_profile.py [preset] [frames]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from groundsegmentationrosnode_b200 import capi, params as P, synth  # noqa: E402

preset = int(sys.argv[1]) if len(sys.argv) > 1 else P.GS_PRESET_HDL64E_2048
F = int(sys.argv[2]) if len(sys.argv) > 2 else 148
p = P.preset(preset)
pool, offs = synth.make_pool(p, 4)
pts, offsets = synth.tile_pool(pool, offs, F)
ctx = capi.Context(p)
ctx.set_option(capi.GS_OPT_PHASE_TIMING, 1)
H, W = ctx.shape
d_pts = torch.from_numpy(pts).cuda()
d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
d_xyz = torch.zeros((F, H * W, 4), dtype=torch.float32, device="cuda")
d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for it in range(2):
    ctx.reset_stats()
    ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
    ctx.synchronize()
ph = ctx.phase_cycles()
st = ctx.stats()
tot = sum(ph.values()) or 1
print("cycles per phase, summed over CTAs")
for k, v in ph.items():
    if v:
        print("  %-26s %14d  %5.1f%%  (%.0f cycles per frame per CTA-equivalent)" % (k, v, 100.0 * v / tot, v / F))
print("fallbacks %d, mismatches %d, remote points %d, dropped %d" % (st.slow_path_points, st.fast_path_mismatches, st.remote_points, st.dropped_points))
