"""CPU tests of the multi-GPU host logic: shard bounds and the world_size-2 gather of limb matrices over gloo
(the GPU path runs the same code over nccl)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pailliercryptolib_python_b200.sharding import gather_rows, shard_bounds, shard_sizes


def test_shard_bounds_cover_and_balance():
    for count in (0, 1, 7, 8, 100000, 8388608, 13):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(count, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(count, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == count
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, counts, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for count in counts:
            rng = np.random.Generator(np.random.PCG64(count))
            full = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=(count, 128), dtype=np.int64).astype(np.int32))
            lo, hi = shard_bounds(count, world, rank)
            got = gather_rows(full[lo:hi].clone(), count)
            assert got.shape == full.shape and torch.equal(got, full), "count=%d rank=%d" % (count, rank)
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_gather_rows_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), [8, 7, 1, 1000], str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
