#!/bin/bash
# racecheck of the default path with every warning printed (after a change to k_decode_sequences' table load)
set -u
mkdir -p gpurun_out
TAG=${1:-r03v}
timeout -s KILL 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(config2_text_frames_small or decodecorpus_batch or dictionaries or config4) and exec2" > gpurun_out/${TAG}_racecheck.txt 2>&1
grep -c "Race reported" gpurun_out/${TAG}_racecheck.txt; grep -v "^=========\s*$" gpurun_out/${TAG}_racecheck.txt | cut -c1-260 | head -40
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | cut -c1-200
