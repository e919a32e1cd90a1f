#!/bin/bash
# k_execute2 (exec2.cuh) against k_execute: tests, the headline workload with build and occupancy variants, the other workloads, ncu
set -u
mkdir -p gpurun_out
TAG=${1:-r02i}
echo "== pytest gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.log
run() {  # name, env..., then bench args after --
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout -s KILL 200 python bench.py "$@" --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${TAG}_$name.err > gpurun_out/${TAG}_$name.json
  python - "$name" gpurun_out/${TAG}_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], "exec %.3f" % d["roofline"]["stages_ms"]["k_execute"], "verified", d["verified"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
L=$PWD/sparkzstd_b200
run text_base A=1 -- 
run text_legacy SZB_EXEC=legacy --
for v in x2c6 x2c4 x2w1 x2w8 x2w2; do run text_$v SZB200_LIB=$L/libszb200_$v.so -- ; done
for n in 2 3 4 6; do run text_cap$n SZB_X2_CTAS_PER_SM=$n -- ; done
run text_w1cap16 SZB200_LIB=$L/libszb200_x2w1.so SZB_X2_CTAS_PER_SM=16 --
run text_w1cap24 SZB200_LIB=$L/libszb200_x2w1.so SZB_X2_CTAS_PER_SM=24 --
run mixed_base A=1 -- --workload mixed
run mixed_legacy SZB_EXEC=legacy -- --workload mixed
run literal_base A=1 -- --workload literal
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_text.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_text.log 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_text.csv
ncu --set full --clock-control none --import-source on -k regex:"k_execute2" -c 1 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
