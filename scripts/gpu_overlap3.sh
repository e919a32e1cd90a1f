#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r03d}
OUT=gpurun_out/${TAG}_overlap3.txt
: > $OUT
run() { env "$@" 2>&1 | tail -1 | tee -a $OUT; }
run SZB_SPLIT=0 timeout -s KILL 300 python scripts/overlap3_exp.py
run SZB_SPLIT=1 timeout -s KILL 300 python scripts/overlap3_exp.py
run SZB_SPLIT=1 SZB_SEQ_CTAS_PER_SM=2 timeout -s KILL 300 python scripts/overlap3_exp.py
run SZB_SPLIT=1 SZB_SEQ_CTAS_PER_SM=2 SZB_X2_CARVEOUT=100 timeout -s KILL 300 python scripts/overlap3_exp.py
run SZB_SPLIT=1 SZB_SEQ_CTAS_PER_SM=1 SZB_X2_CARVEOUT=100 timeout -s KILL 300 python scripts/overlap3_exp.py
run SZB_SPLIT=1 SZB_SEQ_CTAS_PER_SM=3 SZB_X2_CARVEOUT=100 timeout -s KILL 300 python scripts/overlap3_exp.py
