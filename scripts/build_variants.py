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
    # k_execute2 (exec2.cuh): warps per CTA x CTAs per SM the register allocation aims at (launch bounds); the default is 1 x 32
    "x2w4": ("SZB_EXEC2_WARPS=4", "SZB_EXEC2_MIN_CTAS=8"),
    "x2w2": ("SZB_EXEC2_WARPS=2", "SZB_EXEC2_MIN_CTAS=16"),
    "x2w1c24": ("SZB_EXEC2_WARPS=1", "SZB_EXEC2_MIN_CTAS=24"),   # 85 registers
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(b.build(force=True, out=os.path.join(b.HERE, f"libszb200_{n}.so"), defines=VARIANTS[n]))
