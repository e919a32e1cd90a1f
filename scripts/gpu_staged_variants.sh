#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02q}
L=$PWD/sparkzstd_b200
for v in base x2ns x2d4 x2d8 x2d32; do
  lib=$L/libszb200_$v.so; [ "$v" = "base" ] && lib=$L/libszb200.so
  for wl in text mixed; do
  SZB200_LIB=$lib timeout -s KILL 200 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${TAG}_${wl}_$v.err > gpurun_out/${TAG}_${wl}_$v.json
  python - "$v $wl" gpurun_out/${TAG}_${wl}_$v.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], "exec %.3f" % d["roofline"]["stages_ms"]["k_execute"], "verified", d["verified"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
echo "== gpu tests on the staged build"; timeout -s KILL 900 python -m pytest tests -m gpu -q -k exec2 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:"k_execute2" -c 1 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
