#!/bin/bash
# k_execute_team (SZB_PAIR2=2) against k_execute_pair2 (1): forced long-path tests, the mixed corpus, one 64 MiB frame; team sizes
set -u
mkdir -p gpurun_out
TAG=${1:-r03g}
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_frame_paths or config5 or several_groups" 2>&1 | tail -5 | cut -c1-600 | tee gpurun_out/${TAG}_pytest_long.log
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout -s KILL 600 python bench.py "$@" --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - "$name" gpurun_out/${TAG}_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
L=$PWD/sparkzstd_b200
run mixed_team2 SZB_PAIR2=2 -- --workload mixed
run mixed_pair2 SZB_PAIR2=1 -- --workload mixed
run mixed_team4 SZB_PAIR2=2 SZB200_LIB=$L/libszb200_team4.so -- --workload mixed
run mixed_team4c8 SZB_PAIR2=2 SZB200_LIB=$L/libszb200_team4c8.so -- --workload mixed
run mixed_team1 SZB_PAIR2=2 SZB200_LIB=$L/libszb200_team1.so -- --workload mixed
run single64m_team2 SZB_PAIR2=2 SZB_LONG_MODE=pair -- --workload single --frames 1024
run single64m_team4 SZB_PAIR2=2 SZB_LONG_MODE=pair SZB200_LIB=$L/libszb200_team4c8.so -- --workload single --frames 1024
