#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02r}
L=$PWD/sparkzstd_b200
for v in base sq10 sq14 sq16 sq28 x2w2c24 x2w2c20; do
  lib=$L/libszb200_$v.so; [ "$v" = "base" ] && lib=$L/libszb200.so
  SZB200_LIB=$lib timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${TAG}_text_$v.err > gpurun_out/${TAG}_text_$v.json
  python - "$v" gpurun_out/${TAG}_text_$v.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
