#!/bin/bash
# huf_list in family order (longest streams first) against block order: mixed, literal-heavy and text workloads
set -u
mkdir -p gpurun_out
TAG=${1:-r03e}
for p in 1 0; do
  for w in mixed literal text; do
    SZB_LIT_SORT=$p timeout -s KILL 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_${w}_litsort_$p.json 2> gpurun_out/${TAG}_${w}_litsort_$p.err
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*litsort_*.json")):
    try:
        d = json.load(open(f))
        print(f, "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
    except Exception as ex:
        print(f, "FAILED", ex)
PY
