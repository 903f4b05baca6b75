#!/bin/bash
# Run on the GPU box through gpurun: kernel tests, parity tests, smoke, a short bench.  Logs -> gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
K="${VCT_TEST_FILTER-not tcgen05}"
timeout 1200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/test_kernels.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$K" -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/test_parity.log
VCT_GEMM="${VCT_GEMM:-simt}" timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:---gemm simt} > gpurun_out/bench.log 2>&1
echo "== kernels"; tail -25 gpurun_out/test_kernels.log
echo "== parity"; tail -40 gpurun_out/test_parity.log
echo "== smoke"; tail -8 gpurun_out/smoke.log
echo "== bench"; tail -5 gpurun_out/bench.log
