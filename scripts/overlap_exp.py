"""Experiment: stage 3 of one sub-batch beside stage 4 of another (two contexts = two streams), K sub-batches."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from tools import corpus as cg
from sparkzstd_b200.decompression import Context, Batch

N = 65536
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c = cg.config2_text_frames(N)
d_src = torch.from_numpy(c.src).cuda()
D = c.decompressed_bytes
d_dst = torch.empty(D + 512, dtype=torch.uint8, device='cuda')
ctxs = [Context(0), Context(0)]
per = N // K
parts, ptrs, caps = [], [], []
pos = 0
for k in range(K):
    lo, hi = k * per, (k + 1) * per
    parts.append(Batch(ctxs[k % 2], c.src, c.frame_off[lo:hi], c.frame_len[lo:hi]))
    sz = int(c.raw_size[lo:hi].sum())
    ptrs.append(d_dst.data_ptr() + pos)
    caps.append(sz + 256)
    pos += sz

def run_all():
    for k in range(K):
        parts[k].run(d_src.data_ptr(), ptrs[k], caps[k])

for _ in range(3):
    run_all()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    run_all()
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
print(f"K={K} pad={os.environ.get('SZB_K2_PAD','0')}: {dt*1e3:8.2f} ms  {D/dt/1e9:7.1f} GB/s", flush=True)
