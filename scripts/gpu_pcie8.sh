#!/bin/bash
# what the host side of the box moves with N GPUs copying at once, then the default workload's weak-scaling line (value + e2e)
set -u
mkdir -p gpurun_out
TAG=$1; N=$2
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/pcie_bw_multi.py > gpurun_out/${TAG}_pcie_${N}gpu.json 2> gpurun_out/${TAG}_pcie_${N}gpu.err
cat gpurun_out/${TAG}_pcie_${N}gpu.json
nproc; free -g | head -2; lscpu | grep -i "numa\|socket\|model name" | head -8
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_text_weak_${N}gpu.json 2> gpurun_out/${TAG}_text_weak_${N}gpu.err
python - gpurun_out/${TAG}_text_weak_${N}gpu.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e") or {}
    print("GB/s %.1f" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e", e)
except Exception as ex:
    print("FAILED", ex)
PY
