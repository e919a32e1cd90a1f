"""[historical: needs the SZB_SPLIT switch of commits f886271..f25ccb1; results in profiles/r03b-d_*]
Experiment: stage 4 of sub-batch k beside the entropy stages of sub-batch k+1 inside ONE context (SZB_SPLIT=1: stage 4 on a
low-priority stream of its own; SZB_SEQ_CTAS_PER_SM=n: k_decode_sequences capped at n CTAs per SM, so that stage 4's CTAs fit
beside it).  usage: overlap2_exp.py FRAMES_PER_SUBBATCH [FIRST_SUBBATCH_FRAMES]"""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from tools import corpus as cg
from sparkzstd_b200.decompression import Context, Batch

N = int(os.environ.get("N_FRAMES", 65536))
per = int(sys.argv[1]) if len(sys.argv) > 1 else N
first = int(sys.argv[2]) if len(sys.argv) > 2 else per
c = cg.config2_text_frames(N)
d_src = torch.from_numpy(c.src).cuda()
D = c.decompressed_bytes
d_dst = torch.zeros(D + 512, dtype=torch.uint8, device='cuda')
ctx = Context(0)
cuts = [0]
while cuts[-1] < N:
    cuts.append(min(N, cuts[-1] + (first if len(cuts) == 1 else per)))
parts, ptrs, caps, starts = [], [], [], []
pos = 0
for lo, hi in zip(cuts[:-1], cuts[1:]):
    parts.append(Batch(ctx, c.src, c.frame_off[lo:hi], c.frame_len[lo:hi]))
    sz = int(c.raw_size[lo:hi].sum())
    ptrs.append(d_dst.data_ptr() + pos)
    starts.append(pos)
    caps.append(sz + 256)
    pos += sz

def run_all():
    for k in range(len(parts)):
        parts[k].run(d_src.data_ptr(), ptrs[k], caps[k])

for _ in range(3):
    run_all()
torch.cuda.synchronize()
t = time.perf_counter()
R = 5
for _ in range(R):
    run_all()
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / R
# every sub-batch's statuses, and a sample of the frames byte for byte (generator hashes)
ok = True
for p in parts:
    ok = ok and bool((p.finish() == 0).all())
out = d_dst.cpu().numpy()
off = np.concatenate([[0], np.cumsum(c.raw_size)]).astype(np.int64)
for i in range(0, N, 97):
    ok = ok and cg.hash_bytes(out[off[i]:off[i + 1]]) == int(c.raw_hash[i])
print(f"split={os.environ.get('SZB_SPLIT','0')} cap={os.environ.get('SZB_SEQ_CTAS_PER_SM','0')} per={per} first={first} parts={len(parts)}: "
      f"{dt*1e3:8.2f} ms  {D/dt/1e9:7.1f} GB/s verified={ok}", flush=True)
