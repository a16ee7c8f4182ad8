"""Frame-parallel sharding of a batch across the GPUs of one box.

Frames are independent on this path (Alfa_GS::process_pointcloud holds no cross-frame state, reference
src/alfa_gs.cpp:196-239), so rank g of G takes the contiguous block [g*F/G, (g+1)*F/G) and runs it on its own
gs_ctx; there is no data-path collective.  torch.distributed is used only for the timing barrier / max-over-ranks and,
optionally, to gather per-frame point counts (a few bytes per frame)."""
import numpy as np


def shard_range(n_frames, rank, world):
    """Contiguous block of frames owned by `rank`: [lo, hi)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def shard_batch(points, offsets, rank, world):
    """Slice a packed batch (points (sum N, 4) float32, offsets int64 (F + 1)) down to the frames of `rank`.
    Returns (points_view, local_offsets starting at 0, first_frame)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    lo, hi = shard_range(len(offsets) - 1, rank, world)
    p0, p1 = int(offsets[lo]), int(offsets[hi])
    return points[p0:p1], offsets[lo:hi + 1] - p0, lo


def max_over_ranks(value, dist=None, device="cpu"):
    """max of a host scalar over all ranks (the step time of a multi-GPU run is the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_counts(local_counts, dist=None, device="cpu"):
    """Per-frame output point counts of every shard, concatenated in frame order on every rank."""
    local = np.asarray(local_counts, dtype=np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch

    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s[0]) for s in sizes))
    buf = torch.zeros(m, dtype=torch.int64, device=device)
    buf[: local.shape[0]] = torch.from_numpy(local).to(device)
    outs = [torch.zeros(m, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(outs, buf)
    return np.concatenate([o[: int(s[0])].cpu().numpy() for o, s in zip(outs, sizes)])
