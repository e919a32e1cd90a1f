#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02d}
echo "== pytest gpu (place)"; timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.log
scripts/gpu_variants.sh ${TAG}_text --workload text -- base p12 p9 p8
scripts/gpu_variants.sh ${TAG}_mixed --workload mixed -- base
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_text.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_text.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_mixed.csv python bench.py --workload mixed --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_mixed.log 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_text.csv gpurun_out/${TAG}_launches_mixed.csv
echo "== overlap experiment: K sub-batches on two streams"
for K in 1 4 8 16; do python scripts/overlap_exp.py $K 2>&1 | tail -1; done
for K in 1 8; do SZB_EXEC=legacy python scripts/overlap_exp.py $K 2>&1 | tail -1 | sed 's/^/legacy /'; done
ncu --set full --clock-control none --import-source on -k regex:"k_resolve|k_place$" -c 2 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
