#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02g}
echo "== pytest gpu (place)"; timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.log
scripts/gpu_variants.sh ${TAG}_text --workload text -- base r32
scripts/gpu_variants.sh ${TAG}_mixed --workload mixed -- base
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_text.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_text.log 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_text.csv
ncu --set full --clock-control none --import-source on -k regex:"k_resolve" -c 1 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
