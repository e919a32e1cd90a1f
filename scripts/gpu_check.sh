#!/bin/bash
# quick validation on one B200: every GPU test, smoke, the default bench line without the CPU arm
set -u
mkdir -p gpurun_out
TAG=${1:-chk}
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-400 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
for wl in text mixed; do
timeout -s KILL 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err
python - gpurun_out/${TAG}_$wl.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "e2e %.1f" % d["e2e"]["value"], "walk_s %.3f" % d["config"]["header_walk_s"], "verified", d["verified"])
except Exception as e:
    print("FAILED", e)
PY
done
