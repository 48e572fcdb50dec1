"""Multi-GPU plan of the path: reference views (clusters) are independent problems (inference.py:105-119), so
they are partitioned over ranks with no data-path collective; results are gathered on rank 0 (host side)."""
from __future__ import annotations

from typing import List, Sequence


def shard_views(num_views: int, rank: int, world_size: int, mode: str = "round_robin") -> List[int]:
    """Indices of the reference views rank `rank` processes.  Every view is owned by exactly one rank."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    if num_views < 0:
        raise ValueError("num_views must be >= 0")
    if mode == "round_robin":
        return list(range(rank, num_views, world_size))
    if mode == "contiguous":
        base, extra = divmod(num_views, world_size)
        start = rank * base + min(rank, extra)
        return list(range(start, start + base + (1 if rank < extra else 0)))
    raise ValueError(f"unknown sharding mode {mode!r}")


def gather_maps(local_indices: Sequence[int], local_maps, num_views: int, group=None):
    """Collect per-view result tensors on rank 0: returns a list of length num_views (None on other ranks).

    local_maps: list of equally shaped CPU or CUDA tensors, one per entry of local_indices.  Uses
    torch.distributed gather_object-free tensor gathers so it works on gloo (CPU tests) and nccl.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [len(shard_views(num_views, r, world)) for r in range(world)]
    if list(local_indices) != shard_views(num_views, rank, world):
        raise ValueError("local_indices must be this rank's round-robin shard")
    max_n = max(counts) if counts else 0
    if max_n == 0:
        return [] if rank == 0 else None
    ref = local_maps[0] if len(local_maps) else None
    # NCCL moves device tensors only: the shape exchange and the buffers of a rank without views live on its GPU
    on_gpu = dist.get_backend(group) == "nccl"
    device = ref.device if ref is not None else (torch.device("cuda", torch.cuda.current_device()) if on_gpu
                                                 else torch.device("cpu"))
    shape = torch.tensor(list(ref.shape) if ref is not None else [0, 0], dtype=torch.int64, device=device)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    shp = next((tuple(int(v) for v in s) for s in shapes if int(s.sum()) > 0), None)
    buf = torch.zeros((max_n,) + shp, dtype=torch.float32, device=device)
    for i, m in enumerate(local_maps):
        buf[i] = m
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * num_views
    for r in range(world):
        for i, v in enumerate(shard_views(num_views, r, world)):
            out[v] = gathered[r][i].cpu()
    return out
