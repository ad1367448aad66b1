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


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """One process per GPU: run this process on the CPUs of the NUMA node its GPU hangs off, so that
    the pinned host buffers it allocates afterwards (first touch) and its copy threads are local to
    the GPU's PCIe root.  Reads the GPU's PCI address from NVML and the node's CPU list from sysfs;
    does nothing when either is unavailable.  Returns what it found (for the bench's JSON line)."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() \
            else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        sysfs = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(sysfs + "/numa_node").read())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
    except Exception as exc:   # no NVML, no sysfs, containers without the files: stay unbound
        info["error"] = str(exc)[:80]
    return info
