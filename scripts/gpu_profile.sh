#!/bin/bash
# Round profile capture on the GPU box (under gpurun): launch list of the default bench command and one
# ncu --set full capture of the two dominant kernels.  Reports land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01b}
ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 27 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_execute$|k_decode_sequences|k_decode_literals" -s 9 -c 3 \
    -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
