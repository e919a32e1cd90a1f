#!/bin/bash
# stage 4 beside the next sub-batch's entropy stages (scripts/overlap2_exp.py): does the overlap pay?
set -u
mkdir -p gpurun_out
TAG=${1:-r03c}
OUT=gpurun_out/${TAG}_overlap.txt
: > $OUT
run() { env "$@" 2>&1 | tail -1 | tee -a $OUT; }
run SZB_SPLIT=0 SZB_X2_CARVEOUT=100 timeout -s KILL 300 python scripts/overlap2_exp.py 65536
run SZB_SPLIT=1 SZB_X2_CARVEOUT=100 timeout -s KILL 300 python scripts/overlap2_exp.py 12432
run SZB_SPLIT=1 SZB_X2_CARVEOUT=100 SZB_SEQ_CTAS_PER_SM=2 timeout -s KILL 300 python scripts/overlap2_exp.py 12432
run SZB_SPLIT=1 SZB_X2_CARVEOUT=100 SZB_SEQ_CTAS_PER_SM=2 timeout -s KILL 300 python scripts/overlap2_exp.py 6216
run SZB_SPLIT=1 SZB_X2_CARVEOUT=100 SZB_SEQ_CTAS_PER_SM=3 timeout -s KILL 300 python scripts/overlap2_exp.py 9324
run SZB_SPLIT=1 SZB_X2_CARVEOUT=100 SZB_SEQ_CTAS_PER_SM=2 timeout -s KILL 300 python scripts/overlap2_exp.py 24864
