#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02k}
echo "== the two tests that failed in r02j"
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -vv --tb=short -k "(streamed_reader or corrupted_payloads or pageable or lies_about or concatenated) and exec2" 2>&1 | tail -80 | cut -c1-2000 | tee gpurun_out/${TAG}_pytest_two.log
echo "== all gpu tests"
timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-600 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== racecheck (k_execute2)"
SZB_EXEC=exec2 timeout -s KILL 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(config2_text_frames_small or decodecorpus_batch) and exec2" 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/${TAG}_racecheck.txt
echo "== bench text / mixed"
for wl in text mixed; do
timeout -s KILL 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err
python - gpurun_out/${TAG}_$wl.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
except Exception as e:
    print("FAILED", e)
PY
done
