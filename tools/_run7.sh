D=gpurun_out/s21; mkdir -p $D
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $D/b.log 2>&1
python - > $D/e2e_sweep.log 2>&1 <<'PY'
import time, numpy as np, torch
from groundsegmentationrosnode_b200 import capi, params as P, synth
p = P.preset(P.GS_PRESET_HDL64E_2048)
pool, offs = synth.make_pool(p, 64)
F = 1024
pts, offsets = synth.tile_pool(pool, offs, F)
H, W = p.shape
h_pts = torch.from_numpy(pts).pin_memory()
h_res = torch.empty((F, H, W), dtype=torch.int16).pin_memory()
h_xyz = torch.empty((F * 70000, 4), dtype=torch.float32).pin_memory()
oo = np.zeros(F + 1, np.int64)
for chunk in (0, 16, 33, 66, 99, 148, 296):
    with capi.Context(p) as ctx:
        ctx.set_chunk_frames(chunk)
        for _ in range(2):
            ctx.process_batch_host(h_pts.numpy(), offsets, h_res.numpy().view(np.uint16), h_xyz.numpy(), oo)
        t0 = time.perf_counter()
        for _ in range(4):
            ctx.process_batch_host(h_pts.numpy(), offsets, h_res.numpy().view(np.uint16), h_xyz.numpy(), oo)
        dt = (time.perf_counter() - t0) / 4
        print("chunk %4d: %.1f frames/s, %.2f ms per 1024 frames" % (chunk, F / dt, dt * 1e3), flush=True)
PY
