"""Small batch with every frame forced onto the long-frame kernel SZB_PAIR2 selects (run under compute-sanitizer --tool racecheck:
k_execute_pair2 and k_execute_team hand data from warp to warp through shared memory)."""
import hashlib, os, sys
sys.path.insert(0, '.')
import numpy as np
from tools import corpus as cg
from sparkzstd_b200.decompression import Context

assert os.environ.get("SZB_LONG_SEQS") == "1" and os.environ.get("SZB_LONG_MODE") == "pair"
ctx = Context(0)
gold = cg.golden_frames()[:40]
outs = ctx.decode_batch([d for _, d, _, _ in gold])
assert all(hashlib.sha256(o).hexdigest() == sha for o, (_, _, _, sha) in zip(outs, gold))
t = cg.config2_text_frames(24)
outs = ctx.decode_batch([t.frame(i) for i in range(t.nframes)])
assert all(cg.hash_bytes(np.frombuffer(o, dtype="uint8")) == int(t.raw_hash[i]) for i, o in enumerate(outs))
print("race_long ok, SZB_PAIR2 =", os.environ.get("SZB_PAIR2", "1"), "launches", ctx.launch_count())
