#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02t}
echo "== gpu tests with SZB_SEQ=3"; SZB_SEQ=3 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exec2" 2>&1 | tail -6 | cut -c1-500
for m in 1 3; do
  for wl in text mixed literal; do
  SZB_SEQ=$m timeout -s KILL 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_${wl}_seq$m.json 2> gpurun_out/${TAG}_${wl}_seq$m.err
  python - "seq$m $wl" gpurun_out/${TAG}_${wl}_seq$m.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-800:])
PY
  done
done
SZB_SEQ=3 ncu --set full --clock-control none --import-source on -k regex:"k_decode_sequences3" -c 1 -o gpurun_out/${TAG}_seq3_full python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
