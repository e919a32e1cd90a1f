#!/bin/bash
# GPU check of the block-parallel long-frame path (execute_long.cuh): forced-path parity tests, then the two workloads
# that have long frames.  Usage (under gpurun): scripts/gpu_long.sh [tag]
set -u
mkdir -p gpurun_out
TAG=${1:-r01e}
timeout -s KILL 280 python -m pytest tests/test_gpu_parity.py -q -x -k "long_frame_paths" 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_long.log
tail -3 gpurun_out/${TAG}_pytest_long.log
timeout -s KILL 120 python bench.py --workload single --frames 4096 --steps 3 --warmup 3 --no-e2e --no-cpu \
    2> gpurun_out/${TAG}_bench_single256m_1gpu.err > gpurun_out/${TAG}_bench_single256m_1gpu.json
tail -2 gpurun_out/${TAG}_bench_single256m_1gpu.err; cut -c1-300 gpurun_out/${TAG}_bench_single256m_1gpu.json
timeout -s KILL 150 python bench.py --workload mixed --steps 3 --warmup 3 --no-e2e --no-cpu \
    2> gpurun_out/${TAG}_bench_mixed_1gpu.err > gpurun_out/${TAG}_bench_mixed_1gpu.json
tail -2 gpurun_out/${TAG}_bench_mixed_1gpu.err; cut -c1-300 gpurun_out/${TAG}_bench_mixed_1gpu.json
