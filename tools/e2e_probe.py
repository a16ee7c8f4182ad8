import time, numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
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
import ctypes as C
with capi.Context(p) as ctx:
    L = ctx._L
    def run(res, xyz):
        rc = L.gs_process_batch_host(ctx._h, h_pts.numpy().ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), F,
                                     res, xyz, C.c_uint64(h_xyz.shape[0]), oo.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
    pr = h_res.numpy().ctypes.data_as(C.c_void_p); px = h_xyz.numpy().ctypes.data_as(C.c_void_p)
    for name, a, b in (("input only", None, None), ("input + image", pr, None), ("input + image + cloud", pr, px)):
        for _ in range(2): run(a, b)
        t0 = time.perf_counter()
        for _ in range(4): run(a, b)
        dt = (time.perf_counter() - t0) / 4
        print("%-24s %.1f frames/s  (H2D %.1f GB/s)" % (name, F / dt, 16.0 * offsets[-1] / dt / 1e9), flush=True)
# raw pinned H2D / D2H bandwidth
d = torch.empty(h_pts.numel(), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for _ in range(2): d.copy_(h_pts.view(-1), non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(h_pts.view(-1), non_blocking=True)
torch.cuda.synchronize()
print("raw H2D %.1f GB/s" % (5 * h_pts.numel() * 4 / (time.perf_counter() - t0) / 1e9))
