#!/bin/bash
# k_execute_pair2 against k_execute_pair: the forced long-path tests, then the mixed corpus and one 64 MiB frame on both
set -u
mkdir -p gpurun_out
TAG=${1:-r03a}
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_frame_paths or config5 or mixed" 2>&1 | tail -5 | cut -c1-400 | tee gpurun_out/${TAG}_pytest_long.log
for p in 1 0; do
  SZB_PAIR2=$p timeout -s KILL 600 python bench.py --workload mixed --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_mixed_pair2_$p.json 2> gpurun_out/${TAG}_mixed_pair2_$p.err
  SZB_PAIR2=$p SZB_LONG_MODE=pair timeout -s KILL 600 python bench.py --workload single --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${TAG}_single64m_pair2_$p.json 2> gpurun_out/${TAG}_single64m_pair2_$p.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r03a_*pair2_*.json")):
    try:
        d = json.load(open(f))
        print(f, "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
    except Exception as ex:
        print(f, "FAILED", ex)
PY
