#!/bin/bash
# memcheck + racecheck of the default path after a late kernel change (the in-process tests; the subprocess tests: gpu_final.sh)
set -u
mkdir -p gpurun_out
TAG=${1:-r03w}
timeout -s KILL 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(stage_level or config2_text_frames_small or config4 or config5 or ragged or corrupted or decodecorpus_batch or dictionaries or concatenated) and exec2" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/${TAG}_memcheck.txt
timeout -s KILL 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(config2_text_frames_small or decodecorpus_batch or config4) and exec2" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/${TAG}_racecheck.txt
