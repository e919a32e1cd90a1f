#!/bin/bash
# Runs on the GPU box under gpurun: smoke, GPU parity tests, a short bench, an ncu launch list.
# Usage: scripts/gpu_check.sh [bench-frames]
set -u
mkdir -p gpurun_out
FR=${1:-8192}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt; (go version || true) >> gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
if ! grep -q " passed" gpurun_out/pytest_gpu.log || grep -q "failed" gpurun_out/pytest_gpu.log; then
  echo "== pytest gpu (serial table builds)"; SZB200_LIB=$PWD/sparkzstd_b200/libszb200_serial.so timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_serial.log
fi
echo "== bench $FR"; timeout -s KILL 900 python bench.py --frames $FR --steps 3 --warmup 3 2> gpurun_out/bench_small.err | tee gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
