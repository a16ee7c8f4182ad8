"""Small cases for compute-sanitizer (one tool per gpurun call, e.g. `compute-sanitizer --tool memcheck python tools/sanitizer_case.py`):
single frames of every shape, a batch large enough to run the cluster kernel AND its software-cluster twin side by side, the
batched Labeler CCL, the staged kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from groundsegmentationrosnode_b200 import capi, params as P, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

for preset in (P.GS_PRESET_VLP16, P.GS_PRESET_HDL64E_2048, P.GS_PRESET_OS1_128, P.GS_PRESET_CLOUD2RANGE_DEFAULT):
    p = P.preset(preset)
    pts = synth.make_adversarial_frame(p, synth.SEED0 + 3, "collisions") if preset == P.GS_PRESET_VLP16 else synth.make_frame(p, synth.SEED0, shuffle=(preset == P.GS_PRESET_OS1_128))
    want = O.process_frame(p, pts)
    with capi.Context(p) as ctx:
        got = ctx.process_frame(pts)
        ok = all(np.array_equal(got[k], want[k]) for k in ("range", "angle", "no_ground", "xyz"))
        ctx.set_option(capi.GS_OPT_FORCE_STAGED, 1)
        got2 = ctx.process_frame(pts)
        ok2 = all(np.array_equal(got2[k], want[k]) for k in ("range", "angle", "no_ground", "xyz"))
        print(preset, pts.shape[0], "fused ok" if ok else "fused MISMATCH", "staged ok" if ok2 else "staged MISMATCH", ctx.stats().kernel_launches, flush=True)

# hybrid: more frames than co-resident hardware clusters, with the labels on top
p = P.preset(P.GS_PRESET_HDL64E_2048)
H, W = p.shape
pool, offs = synth.make_pool(p, 2)
F = 41
pts, offsets = synth.tile_pool(pool, offs, F)
want = [O.process_frame(p, pool[offs[i]:offs[i + 1]]) for i in range(2)]
beta = float(np.radians(10.0))
wl = [O.label_components(p, w["no_ground"], w["angle"], beta)[0] for w in want]
with capi.Context(p) as ctx:
    d_pts = torch.from_numpy(pts).cuda()
    d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
    d_lab = torch.zeros((F, H, W), dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.process_batch_device_labels(d_pts.data_ptr(), offsets, F, d_res.data_ptr(), 0, d_cnt.data_ptr(), d_lab.data_ptr(), beta)
    ctx.synchronize()
    res = d_res.cpu().numpy().view(np.uint16)
    lab = d_lab.cpu().numpy()
    ok = all(np.array_equal(res[f], want[f % 2]["no_ground"]) and np.array_equal(lab[f], wl[f % 2]) for f in range(F))
    print("batch of %d frames (cluster kernel + software clusters + CCL):" % F, "ok" if ok else "MISMATCH", ctx.stats().kernel_launches, flush=True)
