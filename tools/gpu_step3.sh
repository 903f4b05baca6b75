#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 300 python bench.py --config cfg3 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_cfg3_b.log 2>&1
echo "== cfg3 bf16"; tail -1 gpurun_out/bench_cfg3_b.log | cut -c1-300
timeout 300 python bench.py --config cfg3 --precision bf16x6 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_cfg3_x6.log 2>&1
echo "== cfg3 bf16x6"; tail -1 gpurun_out/bench_cfg3_x6.log | cut -c1-300
timeout 300 python bench.py --config cfg3 --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-rooflines > gpurun_out/bench_cfg3_fp32.log 2>&1
echo "== cfg3 fp32 simt"; tail -1 gpurun_out/bench_cfg3_fp32.log | cut -c1-300
for p in bf16x6 bf16x3 fp32; do
timeout 300 python bench.py --config cfg2 --precision $p --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-rooflines > gpurun_out/bench_cfg2_$p.log 2>&1
echo "== cfg2 $p"; tail -1 gpurun_out/bench_cfg2_$p.log | cut -c1-300
done
