"""Raw pinned-memory PCIe bandwidth of this box: context for bench.py's e2e numbers.

One process:      python tools/pcie_probe.py
One per GPU:      python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                      --master-port 29511 tools/pcie_probe.py
No kernels run: every rank issues cudaMemcpyAsync from / to its own pinned buffers on two streams, all
ranks at the same time (barrier on both sides), and rank 0 prints per-rank and aggregate GB/s for
H2D alone, D2H alone and both directions together.  If the aggregate stops growing with N while
nothing but copies runs, the ceiling is the host (root complexes / memory / IOMMU of the VM), not
the decode pipeline.
"""
import os
import time

import torch

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()


def run(h2d, d2h, reps=6):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    mine = time.perf_counter() - t0
    barrier()
    t = torch.tensor([mine], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    slowest = float(t.item())
    return reps * n / mine / 1e9, world * reps * n / slowest / 1e9


run(True, True, 1)
rows = [("H2D alone", run(True, False)), ("D2H alone", run(False, True)), ("both ways", run(True, True))]
if rank == 0:
    cpus = len(os.sched_getaffinity(0))
    print(f"ranks {world}, host CPUs visible to rank 0: {cpus}, buffers {n >> 20} MiB pinned per direction per rank")
    for name, (mine, agg) in rows:
        print(f"{name}: rank 0 {mine:6.1f} GB/s per direction, all {world} ranks together {agg:7.1f} GB/s per direction")
if dist is not None:
    dist.destroy_process_group()
