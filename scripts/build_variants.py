#!/usr/bin/env python3
"""Builds libszb200_<name>.so for build switches of the kernels, so that one gpurun call can
time them side by side:   python scripts/build_variants.py && gpurun -- 'scripts/gpu_variants.sh TAG --workload single --frames 16384 -- base <names>'
(the .so files are git-ignored but travel to the GPU box).  Every variant decodes the same bytes: tests/test_hostsim.py runs
the switched kernels on the emulated CTA, bench.py verifies every timed run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkzstd_b200 import build as b  # noqa: E402

VARIANTS = {
    # k_decode_sequences: blocks (state chains) per warp; shared memory then fits 228 KB / (lanes x 2.7 KB) one-warp CTAs per SM
    "sq10": ("SZB_SEQ_LANES=10",),   # 8 CTAs per SM: two warps per scheduler, half the lanes busy
    "sq14": ("SZB_SEQ_LANES=14",),   # 6 CTAs per SM
    "sq16": ("SZB_SEQ_LANES=16",),   # 5 CTAs per SM
    "sq28": ("SZB_SEQ_LANES=28",),   # 3 CTAs per SM
    # k_execute2 (exec2.cuh): rounds of short sequences staged in shared memory, one lane per sequence
    "x2st": ("SZB_X2_STAGED=1",),
    # more than 32 warps per SM need CTAs of two warps (32 CTAs per SM is the limit) and fewer registers per thread
    "x2w2c24": ("SZB_EXEC2_WARPS=2", "SZB_EXEC2_MIN_CTAS=24"),   # 48 warps per SM, 42 registers
    "x2w2c20": ("SZB_EXEC2_WARPS=2", "SZB_EXEC2_MIN_CTAS=20"),   # 40 warps per SM, 51 registers
    # k_execute_team (exec2.cuh): consumer warps per long frame (default 2); 11 CTAs per SM hold the mixed corpus' 1 630 long frames
    "team4": ("SZB_X2_TEAM=4", "SZB_TEAM_MIN_CTAS=11"),   # 160 threads x 11 CTAs: 37 registers
    "team4c8": ("SZB_X2_TEAM=4", "SZB_TEAM_MIN_CTAS=8"),  # 51 registers, 1 184 frames resident
    "team1": ("SZB_X2_TEAM=1",),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(b.build(force=True, out=os.path.join(b.HERE, f"libszb200_{n}.so"), defines=VARIANTS[n]))
