"""Multi-GPU host logic on CPU: two gloo ranks shard a batch by image, broadcast the
tables once, decode their shard with the CPU oracle and agree with the unsharded result."""
import os
import socket
import sys

import numpy as np
import pytest

import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_contiguous_shards_cover_everything():
    for n in (1, 7, 256, 1024):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_contiguous(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_lpt_balances_mixed_batch():
    rng = np.random.default_rng(0)
    costs = [int(c) for c in rng.choice([512 * 512, 1920 * 1080, 3840 * 2160], size=64)]
    parts = shard.shard_lpt(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(64))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= 1.25 * (sum(costs) / 8)
    assert parts == shard.shard_lpt(costs, 8)   # deterministic: every rank computes the same map


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle
    from util import make_batch, oracle_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = [(64, 48, "420"), (40, 24, "444"), (56, 24, "422"), (48, 40, "gray"), (70, 50, "420"), (33, 17, "422")]
    lo, hi = shard.shard_contiguous(len(shapes), world, rank)
    # rank 0 owns the tables; the others start with garbage and must end up with rank 0's
    q = synth.quality_tables(85) if rank == 0 else np.full((4, 64), 7, dtype=np.uint16)
    q = shard.broadcast_tables(q)
    descs, coef_len, rgb_len, _ = make_batch(shapes[lo:hi], want_yuv=False)
    coef = synth.batch_coefficients(descs, coef_len, q, first_index=lo)
    rgb, _ = oracle_batch(oracle.port(), descs, coef, q, rgb_len, 0, nthreads=1)
    total = shard.sum_over_ranks(float(rgb.astype(np.int64).sum()))
    slow = shard.max_over_ranks(float(rank + 1))
    np.save(out + f".{rank}.npy", rgb)
    if rank == 0:
        np.save(out + ".meta.npy", np.array([total, slow]))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from util import make_batch, oracle_batch
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rgb")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    shapes = [(64, 48, "420"), (40, 24, "444"), (56, 24, "422"), (48, 40, "gray"), (70, 50, "420"), (33, 17, "422")]
    q = synth.quality_tables(85)
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    coef = synth.batch_coefficients(descs, coef_len, q)
    want, _ = oracle_batch(oracle.port(), descs, coef, q, rgb_len, 0)
    got = [np.load(out + f".{r}.npy") for r in range(2)]
    # each rank packed its shard from offset 0: compare image by image
    off = [0, 0]
    for i, d in enumerate(descs):
        r = 0 if i < 3 else 1
        n = d.query_layout().rgb_len
        assert np.array_equal(got[r][off[r]:off[r] + n], want[d.rgb_off:d.rgb_off + n]), i
        off[r] += -(-n // 256) * 256
    meta = np.load(out + ".meta.npy")
    assert meta[0] == float(sum(g.astype(np.int64).sum() for g in got)) and meta[1] == 2.0
