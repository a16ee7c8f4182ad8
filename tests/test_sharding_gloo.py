"""The N > 1 path on CPU: world_size-2 `gloo` processes shard a batch frame-parallel, each processes its block
(with the oracle standing in for the GPU — this test is about the host-side sharding logic), and the gathered result
must equal the single-process result ("N shards == 1 shard")."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from groundsegmentationrosnode_b200 import params as P, sharding, synth  # noqa: E402


def test_shard_ranges_partition_the_batch():
    for F in (0, 1, 7, 64, 4096, 4099):
        for G in (1, 2, 3, 4, 8):
            got = [sharding.shard_range(F, g, G) for g in range(G)]
            assert got[0][0] == 0 and got[-1][1] == F
            for a, b in zip(got, got[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O

    p = P.preset(P.GS_PRESET_VLP16)
    pool, offs = synth.make_pool(p, 3, seed0=synth.SEED0 + 70)
    pts, offsets = synth.tile_pool(pool, offs, 7)  # 7 frames over 2 ranks: 3 + 4
    my_pts, my_offs, first = sharding.shard_batch(pts, offsets, rank, world)
    res = O.process_batch(p, my_pts, my_offs, n_threads=1)
    counts = sharding.gather_counts(res["counts"], dist)
    slowest = sharding.max_over_ranks(10.0 + rank, dist)
    dist.barrier()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), no_ground=res["no_ground"], counts=counts, first=first, slowest=slowest)
    dist.destroy_process_group()


def test_two_shards_equal_one_shard(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    p = P.preset(P.GS_PRESET_VLP16)
    pool, offs = synth.make_pool(p, 3, seed0=synth.SEED0 + 70)
    pts, offsets = synth.tile_pool(pool, offs, 7)
    whole = oracle.process_batch(p, pts, offsets, n_threads=1)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(x["first"]) for x in parts] == [0, 3]
    assert np.array_equal(np.concatenate([x["no_ground"] for x in parts], 0), whole["no_ground"])
    for x in parts:
        assert np.array_equal(x["counts"], whole["counts"].astype(np.int64))
        assert float(x["slowest"]) == 11.0
