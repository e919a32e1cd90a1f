#!/bin/bash
# round 2, the build with k_execute2 as the default: tests on all stage-4 paths, bench lines, long-frame slices, ncu, sanitizer
set -u
mkdir -p gpurun_out
TAG=${1:-r02j}
echo "== pytest gpu"; timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== bench default (text, e2e, cpu baseline)"
timeout -s KILL 600 python bench.py > gpurun_out/${TAG}_bench_text.json 2> gpurun_out/${TAG}_bench_text.err; tail -c 600 gpurun_out/${TAG}_bench_text.json; echo
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout -s KILL 400 python bench.py "$@" --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${TAG}_$name.err > gpurun_out/${TAG}_$name.json
  python - "$name" gpurun_out/${TAG}_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run mixed A=1 -- --workload mixed
run literal A=1 -- --workload literal
run single1g A=1 -- --workload single --frames 16384
run single1g_s1024 SZB_LONG_SLICE=1024 -- --workload single --frames 16384
run single1g_s4096 SZB_LONG_SLICE=4096 -- --workload single --frames 16384
run single64m A=1 -- --workload single --frames 1024
run single64m_s512 SZB_LONG_SLICE=512 -- --workload single --frames 1024
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_text.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_text.log 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_text.csv | tee gpurun_out/${TAG}_launch_shares.txt
ncu --set full --clock-control none --import-source on -k regex:"k_execute2|k_decode_sequences|k_decode_literals|k_build" -c 5 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
echo "== memcheck"
timeout -s KILL 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_frame_paths or config2_text_frames_small or ragged or corrupted or decodecorpus_batch" 2>&1 | tail -8 | tee gpurun_out/${TAG}_memcheck.txt
echo "== racecheck"
timeout -s KILL 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config2_text_frames_small or decodecorpus_batch" 2>&1 | tail -8 | tee gpurun_out/${TAG}_racecheck.txt
