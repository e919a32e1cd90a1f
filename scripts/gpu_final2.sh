#!/bin/bash
# after the last kernel change of the round (k_decode_sequences back to one group per CTA): tests, bench lines, launch list, full capture
set -u
mkdir -p gpurun_out
TAG=${1:-r03y}
echo "== pytest gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-400 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/${TAG}_smoke.log
echo "== bench default"; timeout -s KILL 600 python bench.py > gpurun_out/${TAG}_bench_text.json 2> gpurun_out/${TAG}_bench_text.err; cut -c1-600 gpurun_out/${TAG}_bench_text.json
for w in mixed literal; do
  timeout -s KILL 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; cut -c1-300 gpurun_out/${TAG}_bench_$w.json; echo
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_text.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_text.log 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_text.csv | tee gpurun_out/${TAG}_launch_shares.txt
ncu --set full --clock-control none --import-source on -k regex:"k_execute2|k_decode_sequences|k_decode_literals|k_build" -c 5 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
