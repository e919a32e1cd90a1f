#!/bin/bash
# Everything that was implemented after the GPU budget of round 1 ran out, timed in ONE gpurun call (~4 GPU-minutes):
#   python scripts/build_variants.py && gpurun --timeout 600 -- 'scripts/gpu_pending.sh r02a'
# 1. the default build on one 256 MiB / 1 GiB frame (k_long_jump with batched loads, literal prefetch in k_long_emit),
#    against the loop the round-1 numbers were taken with (variant "unbatched") and the lane-refill loops;
# 2. SZB_LONG_SLICE on the 256 MiB frame (one warp per 1 024 sequences in k_long_hist / k_long_emit);
# 3. the GPU test suite.
# Results: gpurun_out/${TAG}_*.json (bench lines, verified against the generator's hashes).
set -u
mkdir -p gpurun_out
TAG=${1:-r02a}
echo "== variants, one 256 MiB frame"; scripts/gpu_variants.sh ${TAG}_256m --workload single --frames 4096 -- base unbatched refill4 refill8 batched4
echo "== variants, one 1 GiB frame"; scripts/gpu_variants.sh ${TAG}_1g --workload single --frames 16384 -- base unbatched refill4
for S in 512 1024 4096; do
  echo "== SZB_LONG_SLICE=$S, one 256 MiB frame"
  SZB_LONG_SLICE=$S timeout -s KILL 120 python bench.py --workload single --frames 4096 --steps 3 --warmup 3 --no-e2e --no-cpu \
      2> gpurun_out/${TAG}_slice${S}.err > gpurun_out/${TAG}_slice${S}.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('GB/s %.2f ms %.3f exec %.3f verified %s' % (d['value'], d['ms_per_step'], d['roofline']['stages_ms']['k_execute'], d['verified']))" gpurun_out/${TAG}_slice${S}.json
done
echo "== pytest gpu"; timeout -s KILL 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.log
