#!/bin/bash
# multi-GPU lines: configs[4] as BASELINE.json words it (ONE mixed corpus, frame-sharded over the ranks: strong scaling), and
# the default workload (weak scaling) with its end-to-end number.  usage (under gpurun --gpus N): scripts/gpu_multi.sh TAG N [TOTAL_BYTES]
set -u
mkdir -p gpurun_out
TAG=$1; N=$2; TOTAL=${3:-17179869184}
run() {
  local name=$1; shift
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
     > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err
  python - gpurun_out/${TAG}_${name}_${N}gpu.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e") or {}
    print(sys.argv[1], "GB/s %.1f" % d["value"], "ms %.2f" % d["ms_per_step"], "scaling", d["scaling"], "e2e", round(e.get("value", 0), 1), "shards", d["config"].get("shards"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
run mixed_strong --workload mixed --scaling strong --total-bytes $TOTAL --steps 3 --warmup 3 --no-cpu
run text_weak --steps 3 --warmup 3 --no-cpu
