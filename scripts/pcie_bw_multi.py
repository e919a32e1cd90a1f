#!/usr/bin/env python3
"""What the host side of a box moves when every GPU copies at once: one process per GPU (torchrun), pinned buffers, D2H alone,
H2D alone, both directions together; aggregate over the ranks = max-over-ranks time.  The end-to-end bench number of N ranks cannot
exceed this.  usage: python -m torch.distributed.run --nproc-per-node N scripts/pcie_bw_multi.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 2 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


def both():
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)


out = {"n_gpus": world, "bytes_per_copy": n}
out["d2h_GBps_aggregate"] = world * n / timed(lambda: h1.copy_(d1, non_blocking=True)) / 1e9
out["h2d_GBps_aggregate"] = world * n / timed(lambda: d2.copy_(h2, non_blocking=True)) / 1e9
out["bidir_each_direction_GBps_aggregate"] = world * n / timed(both) / 1e9
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
