#!/bin/bash
# Round-end GPU check: full GPU test suite, the default bench line, configs[2] at full size, the mixed corpus, smoke.
set -u
mkdir -p gpurun_out
TAG=${1:-r01k}
echo "== pytest gpu"; timeout -s KILL 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log | tail -3
echo "== text"; timeout -s KILL 200 python bench.py --no-cpu 2> gpurun_out/${TAG}_bench_text_1gpu.err > gpurun_out/${TAG}_bench_text_1gpu.json; cut -c1-160 gpurun_out/${TAG}_bench_text_1gpu.json
echo "== single 4 GiB"; timeout -s KILL 300 python bench.py --workload single --frames 65536 --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${TAG}_bench_single4g_1gpu.err > gpurun_out/${TAG}_bench_single4g_1gpu.json; cut -c1-160 gpurun_out/${TAG}_bench_single4g_1gpu.json
echo "== mixed"; timeout -s KILL 150 python bench.py --workload mixed --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${TAG}_bench_mixed_1gpu.err > gpurun_out/${TAG}_bench_mixed_1gpu.json; cut -c1-160 gpurun_out/${TAG}_bench_mixed_1gpu.json
echo "== smoke"; timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
