"""Multi-GPU host logic: images are independent, so a batch shards by image
with no data-path collective.  The only shared data are the
quantisation tables and the batch description, broadcast once from rank 0.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_contiguous(n_images: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Equal contiguous ranges: rank r owns [r*n/R, (r+1)*n/R)."""
    lo = rank * n_images // world_size
    hi = (rank + 1) * n_images // world_size
    return lo, hi


def shard_lpt(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment for mixed-size batches: images
    sorted by cost (bytes or pixels), each to the currently lightest rank.
    Deterministic (ties broken by index), so every rank computes the same map."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    for lst in out:
        lst.sort()
    return out


def broadcast_tables(qtabs: np.ndarray, device=None, src: int = 0) -> np.ndarray:
    """One broadcast of the (n_sets,4,64) uint16 tables from rank `src` over the
    default process group (NCCL on GPUs, gloo on CPU).  Returns the tables every
    rank must use."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return qtabs
    t = torch.as_tensor(qtabs.astype(np.int32))
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy().astype(np.uint16)


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction: the slowest rank defines the step time."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
