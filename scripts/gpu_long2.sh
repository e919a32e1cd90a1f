#!/bin/bash
# Quick GPU loop for the long-frame path: forced-path parity, the single-frame bench, per-kernel ncu times.
set -u
mkdir -p gpurun_out
TAG=${1:-r01f}
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -q -x -k "long_frame_paths and jump" 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_long.log
tail -3 gpurun_out/${TAG}_pytest_long.log
timeout -s KILL 120 python bench.py --workload single --frames 4096 --steps 3 --warmup 3 --no-e2e --no-cpu \
    2> gpurun_out/${TAG}_bench_single256m_1gpu.err > gpurun_out/${TAG}_bench_single256m_1gpu.json
tail -2 gpurun_out/${TAG}_bench_single256m_1gpu.err; cut -c1-200 gpurun_out/${TAG}_bench_single256m_1gpu.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,lts__t_sectors.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_long" -c 5 --csv \
    --log-file gpurun_out/${TAG}_long_launches.csv python bench.py --workload single --frames 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify > gpurun_out/${TAG}_long_launches.log 2>&1
grep -h "gpu__time_duration" gpurun_out/${TAG}_long_launches.csv | cut -d, -f5,12- | head
