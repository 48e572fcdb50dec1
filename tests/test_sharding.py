"""Host-side multi-GPU logic on CPU: view sharding and the rank-0 gather over a world_size-2 gloo group."""
import os
import socket

import pytest

from mvsnet_b200 import sharding


def test_shards_partition_the_views():
    for n in (0, 1, 7, 49):
        for world in (1, 2, 4, 8):
            for mode in ("round_robin", "contiguous"):
                shards = [sharding.shard_views(n, r, world, mode) for r in range(world)]
                flat = sorted(v for s in shards for v in s)
                assert flat == list(range(n)), (n, world, mode)
                assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        sharding.shard_views(4, 2, 2)
    with pytest.raises(ValueError):
        sharding.shard_views(4, 0, 2, "zigzag")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_views, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_views(num_views, rank, world)
        maps = [torch.full((3, 4), float(v)) for v in mine]          # stand-in for depth maps of view v
        out = sharding.gather_maps(mine, maps, num_views)
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ok = all(o is not None and float(o[0, 0]) == float(v) for v, o in enumerate(out))
            ret.put((ok, len(out), float(t[0])))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_views", [5, 2, 1])
def test_gather_over_gloo_world2(num_views):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_views, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ok, n, tmax = q.get(timeout=10)
    assert ok and n == num_views and tmax == 11.0
