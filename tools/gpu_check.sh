#!/bin/bash
# Run on the GPU box through gpurun: every GPU test, smoke, a short bench.  Logs -> gpurun_out/.
#   gpurun --timeout 600 -- 'bash tools/gpu_check.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider ${VCT_TEST_FILTER:+-k "$VCT_TEST_FILTER"} 2>&1 | tail -40 > gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py ${BENCH_ARGS:---steps 50 --warmup 5} > gpurun_out/bench.log 2>&1
echo "== tests"; tail -6 gpurun_out/tests_gpu.log
echo "== smoke"; tail -6 gpurun_out/smoke.log
echo "== bench"; tail -2 gpurun_out/bench.log | cut -c1-600
