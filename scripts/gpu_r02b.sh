#!/bin/bash
# new stage 4 (k_resolve + k_place) against the legacy k_execute: tests, then the headline workload both ways
set -u
mkdir -p gpurun_out
TAG=${1:-r02b}
echo "== pytest gpu (place)"; timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== pytest gpu (legacy)"; SZB_EXEC=legacy timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu_legacy.log
for mode in place legacy; do
  echo "== text, SZB_EXEC=$mode"
  SZB_EXEC=$mode timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_text_$mode.json 2> gpurun_out/${TAG}_text_$mode.err
  python - gpurun_out/${TAG}_text_$mode.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("GB/s %.2f ms %.3f verified %s" % (d["value"], d["ms_per_step"], d["verified"]), d["roofline"]["stages_ms"], d["config"].get("last_step_ms"))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-2000:])
PY
done
for wl in mixed literal; do
  echo "== $wl"
  timeout -s KILL 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err
  python - gpurun_out/${TAG}_$wl.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("GB/s %.2f ms %.3f verified %s" % (d["value"], d["ms_per_step"], d["verified"]), d["roofline"]["stages_ms"], d["config"].get("last_step_ms"))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-2000:])
PY
done
