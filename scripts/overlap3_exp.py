"""[historical: needs the SZB_SPLIT switch of commits f886271..f25ccb1; results in profiles/r03b-d_*]
Diagnostic: do the entropy stages of batch B run BESIDE stage 4 of batch A (SZB_SPLIT=1)?  Times entropy(B) alone, execute(A) alone
and both launched together."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from tools import corpus as cg
from sparkzstd_b200.decompression import Context, Batch

N = int(os.environ.get("N_FRAMES", 65536))
H = N // 2
c = cg.config2_text_frames(N)
d_src = torch.from_numpy(c.src).cuda()
D = c.decompressed_bytes
d_dst = torch.zeros(D + 512, dtype=torch.uint8, device='cuda')
ctx = Context(0)
A = Batch(ctx, c.src, c.frame_off[:H], c.frame_len[:H])
B = Batch(ctx, c.src, c.frame_off[H:], c.frame_len[H:])
szA = int(c.raw_size[:H].sum())
pA, pB = d_dst.data_ptr(), d_dst.data_ptr() + szA
capA, capB = szA + 256, D - szA + 256
sync = torch.cuda.synchronize

def timed(fn, reps=5):
    fn(); sync()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
        sync()
    return (time.perf_counter() - t) / reps * 1e3

A.decode_entropy(d_src.data_ptr()); B.decode_entropy(d_src.data_ptr()); sync()
t_ent = timed(lambda: B.decode_entropy(d_src.data_ptr()))
t_exe = timed(lambda: A.execute(d_src.data_ptr(), pA, capA))
def both():
    A.execute(d_src.data_ptr(), pA, capA)
    B.decode_entropy(d_src.data_ptr())
t_both = timed(both)
def both2():
    B.decode_entropy(d_src.data_ptr())
    A.execute(d_src.data_ptr(), pA, capA)
t_both2 = timed(both2)
print(f"split={os.environ.get('SZB_SPLIT','0')} cap={os.environ.get('SZB_SEQ_CTAS_PER_SM','0')} carve={os.environ.get('SZB_X2_CARVEOUT','-')} "
      f"half batches: entropy(B) {t_ent:.2f} ms, execute(A) {t_exe:.2f} ms, execute(A) then entropy(B) {t_both:.2f} ms, entropy(B) then execute(A) {t_both2:.2f} ms", flush=True)
