#!/bin/bash
# the round's closing run on one B200: every GPU test, the bench lines of all workloads, launch list, ncu --set full, sanitizers
set -u
mkdir -p gpurun_out
TAG=${1:-r03z}
echo "== pytest gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-400 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
echo "== bench default"; timeout -s KILL 600 python bench.py > gpurun_out/${TAG}_bench_text.json 2> gpurun_out/${TAG}_bench_text.err
echo "== bench reference arm"; timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_text_reference.json 2> gpurun_out/${TAG}_bench_text_reference.err
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout -s KILL 600 python bench.py "$@" --steps 3 --warmup 3 --no-cpu 2> gpurun_out/${TAG}_bench_$name.err > gpurun_out/${TAG}_bench_$name.json
  python - "$name" gpurun_out/${TAG}_bench_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    e = d.get("e2e") or {}
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "pipe %.3f" % d["roofline"]["pipeline"]["frac"], "e2e %.1f" % e.get("value", 0), "verified", d["verified"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
run text A=1 --
run mixed A=1 -- --workload mixed
run literal A=1 -- --workload literal
run single256m A=1 -- --workload single --frames 4096 --no-e2e
run single4g A=1 -- --workload single --frames 65536 --no-e2e
run single4g_s1024 SZB_LONG_SLICE=1024 -- --workload single --frames 65536 --no-e2e
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_text.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_text.log 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_text.csv | tee gpurun_out/${TAG}_launch_shares.txt
ncu --set full --clock-control none --import-source on -k regex:"k_execute2|k_decode_sequences|k_decode_literals|k_build" -c 5 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
echo "== memcheck of the long-frame paths (subprocess tests: --target-processes all)"
SZB_TEST_TIMEOUT=1500 timeout -s KILL 1500 compute-sanitizer --tool memcheck --target-processes all python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(long_frame_paths or several_groups) and exec2" 2>&1 | grep -v "^=========     \|^$" | tail -12 | cut -c1-300 | tee gpurun_out/${TAG}_memcheck_long.txt
echo "== memcheck"
timeout -s KILL 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(long_frame_paths or config2_text_frames_small or ragged or corrupted or decodecorpus_batch or dictionaries or concatenated or config5) and not place" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/${TAG}_memcheck.txt
echo "== racecheck"
timeout -s KILL 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(config2_text_frames_small or decodecorpus_batch or dictionaries) and exec2" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/${TAG}_racecheck.txt
echo "== racecheck of the two-warp / team kernels (every frame forced onto them)"
for p in 1 2; do
  SZB_LONG_SEQS=1 SZB_LONG_MODE=pair SZB_PAIR2=$p timeout -s KILL 900 compute-sanitizer --tool racecheck python scripts/race_long.py 2>&1 | grep -v "^=========     \|^$" | tail -6 | cut -c1-300 | tee -a gpurun_out/${TAG}_racecheck_long.txt
done
