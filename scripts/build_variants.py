#!/usr/bin/env python3
"""Builds libszb200_<name>.so for the build switches that are implemented but not measured yet, so that one gpurun call can
time them side by side:   python scripts/build_variants.py && gpurun -- 'scripts/gpu_variants.sh TAG --workload single --frames 16384 -- base <names>'
(the .so files are git-ignored but travel to the GPU box).  Every variant decodes the same bytes: tests/test_hostsim.py runs
the switched kernels on the emulated CTA, bench.py verifies every timed run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkzstd_b200 import build as b  # noqa: E402

VARIANTS = {
    # k_long_jump as measured in round 1 (profiles/README.md, r01m): a branch per walk, 8 CTAs per SM
    "unbatched": ("SZB_JUMP_BATCHED=0", "SZB_JUMP_CTAS_PER_SM=8"),
    # lanes refill a finished walk with their next byte; 4 walks per lane, tiles of 512 cells
    "refill4": ("SZB_JUMP_REFILL=1", "SZB_JUMP_CHAINS=4", "SZB_JUMP_CTAS_PER_SM=6"),
    # the same with 8 walks per lane: needs tiles of 1 024 cells and 56 registers
    "refill8": ("SZB_JUMP_REFILL=1", "SZB_JUMP_CHAINS=8", "SZB_JUMP_TILE=1024", "SZB_JUMP_CTAS_PER_SM=4"),
    # 4 walks per thread, batched, 6 CTAs per SM
    "batched4": ("SZB_JUMP_CHAINS=4", "SZB_JUMP_CTAS_PER_SM=6"),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(b.build(force=True, out=os.path.join(b.HERE, f"libszb200_{n}.so"), defines=VARIANTS[n]))
