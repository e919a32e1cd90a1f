#!/bin/bash
# Times build variants of libszb200 (experiment switches, -D...) on one workload.  Usage: scripts/gpu_variants.sh tag workload-args -- variant...
set -u
mkdir -p gpurun_out
TAG=$1; shift
ARGS=()
while [ "$1" != "--" ]; do ARGS+=("$1"); shift; done
shift
for v in "$@"; do
  lib=$PWD/sparkzstd_b200/libszb200_$v.so
  [ "$v" = "base" ] && lib=$PWD/sparkzstd_b200/libszb200.so
  SZB200_LIB=$lib timeout -s KILL 120 python bench.py "${ARGS[@]}" --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${TAG}_$v.err > gpurun_out/${TAG}_$v.json
  python - "$v" gpurun_out/${TAG}_$v.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], "exec %.3f" % d["roofline"]["stages_ms"]["k_execute"], "verified", d["verified"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
