#!/bin/bash
# all GPU tests, then the text / literal / mixed bench lines (device-timed only)
set -u
mkdir -p gpurun_out
TAG=${1:-r03h}
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | cut -c1-600 | tee gpurun_out/${TAG}_pytest_gpu.log
for w in text literal mixed; do
  timeout -s KILL 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  python - $w gpurun_out/${TAG}_bench_$w.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_literal.csv python bench.py --workload literal --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_literal.log 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_literal.csv | tee gpurun_out/${TAG}_launch_shares_literal.txt
