#!/bin/bash
# k_decode_literals with an L2 prefetch of the stream below the ring's requests (SZB_HUF_PF_BYTES): literal, mixed, text
set -u
mkdir -p gpurun_out
TAG=${1:-r03i}
L=$PWD/sparkzstd_b200
for v in base hpf512 hpf1k hpf2k hpf4k; do
  lib=$L/libszb200.so; [ $v != base ] && lib=$L/libszb200_$v.so
  for w in literal mixed text; do
    SZB200_LIB=$lib timeout -s KILL 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_${w}_$v.json 2> gpurun_out/${TAG}_${w}_$v.err
    python - $v-$w gpurun_out/${TAG}_${w}_$v.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "GB/s %.2f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}, "verified", d["verified"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
  done
done 2>&1 | tee gpurun_out/${TAG}_hpf.txt
