"""Multi-GPU host split: world_size-2 gloo run on CPU + properties of the binning."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_a_balanced_exact_cover():
    from sparkzstd_b200.sharding import shard_frames

    rng = np.random.default_rng(1)
    w = rng.integers(1, 4 << 20, size=1000)
    for world in (1, 2, 4, 8):
        parts = shard_frames(w, world)
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(1000))
        loads = [int(w[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= int(w.max())
    assert [len(p) for p in shard_frames([], 4)] == [0, 0, 0, 0]
    assert [p.tolist() for p in shard_frames([5], 2)] == [[0], []]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from sparkzstd_b200.sharding import agree_on_partition

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 1 deliberately starts from different weights: the broadcast must win
    rng = np.random.default_rng(0 if rank == 0 else 99)
    w = rng.integers(1, 1 << 20, size=257)
    mine = agree_on_partition(w, world, rank)
    import torch

    total = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(total)
    t = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks timing reduction
    q.put((rank, mine.tolist(), int(total.item()), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_partition_agreement():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, idx0, tot0, t0), (r1, idx1, tot1, t1) = res
    assert tot0 == tot1 == 257 and sorted(idx0 + idx1) == list(range(257))
    assert t0 == t1 == 0.002
    from sparkzstd_b200.sharding import shard_frames

    w = np.random.default_rng(0).integers(1, 1 << 20, size=257)
    parts = shard_frames(w, 2)
    assert idx0 == parts[0].tolist() and idx1 == parts[1].tolist()
