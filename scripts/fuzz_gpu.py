#!/usr/bin/env python3
"""One-off robustness run on the GPU box: corrupted frames (bit flips, byte stomps, truncations) through the C ABI.
The engine must return for every batch; where the oracle still decodes a frame and the engine reports success the bytes
must be identical; frames the oracle rejects but the engine accepts are listed (the engine may only be stricter).
usage: scripts/fuzz_gpu.py [variants] [seed]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from oracle import pyszo  # noqa: E402
from tools import corpus as cg  # noqa: E402
import crafted_frames  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    base = [d for _, d, _, _ in cg.golden_frames() if len(d) > 40]
    t = cg.config2_text_frames(16)
    base += [t.frame(i) for i in range(t.nframes)]
    m = cg.config5_mixed(8 << 20, with_golden=False)
    base += [m.frame(i) for i in range(m.nframes) if m.frame_len[i] < (1 << 20)]
    base += [f for f, _ in crafted_frames.cases().values()]
    frames = []
    for _ in range(n):
        buf = bytearray(base[int(rng.integers(0, len(base)))])
        kind = int(rng.integers(0, 4))
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                buf[int(rng.integers(4, len(buf)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            p = int(rng.integers(4, len(buf)))
            for k in range(p, min(len(buf), p + int(rng.integers(1, 9)))):
                buf[k] = int(rng.integers(0, 256))
        elif kind == 2:
            del buf[int(rng.integers(6, len(buf))):]
        else:
            p = int(rng.integers(5, min(len(buf), 64)))
            buf[p] = int(rng.integers(0, 256))
        frames.append(bytes(buf))

    from sparkzstd_b200.decompression import Context

    ctx = Context(0)
    stats = {"variants": n, "both_ok": 0, "both_fail": 0, "engine_stricter": 0, "engine_laxer": [], "mismatch": []}
    dst = np.empty(512 << 20, dtype=np.uint8)
    for lo in range(0, n, 250):
        chunk = frames[lo : lo + 250]
        src = np.frombuffer(b"".join(chunk) + b"\0" * 16, dtype=np.uint8)
        lens = np.array([len(f) for f in chunk], dtype=np.uint64)
        offs = (np.cumsum(lens) - lens).astype(np.uint64)
        try:
            out_off, out_len, status = ctx.decode_batch_into(src, offs, lens, dst)
        except Exception as e:  # the whole call failing is legal only for capacity reasons
            print("batch", lo, "raised", repr(e)[:200])
            stats.setdefault("batch_errors", []).append(lo)
            continue
        for i, f in enumerate(chunk):
            try:
                want = pyszo.decode_frame(f)
            except pyszo.OracleError:
                want = None
            ok = status[i] == 0
            if want is not None and ok:
                o, l = int(out_off[i]), int(out_len[i])
                if dst[o : o + l].tobytes() == want:
                    stats["both_ok"] += 1
                else:
                    stats["mismatch"].append(lo + i)
            elif want is None and not ok:
                stats["both_fail"] += 1
            elif want is not None:
                stats["engine_stricter"] += 1
                key = str(int(status[i]))
                stats.setdefault("stricter_by_status", {})[key] = stats.setdefault("stricter_by_status", {}).get(key, 0) + 1
            else:
                stats["engine_laxer"].append(lo + i)
    ctx.close()
    print(json.dumps(stats))
    return 1 if stats["mismatch"] else 0


if __name__ == "__main__":
    sys.exit(main())
