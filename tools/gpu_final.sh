#!/bin/bash
# Round-end verification on one B200: every GPU test, smoke, the default bench line (with both reference arms), the other
# BASELINE configs, then the ncu launch list + --set full summaries of the current kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG="${PROFILE_TAG:-r02d}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/tests_gpu.log
echo "== tests"; tail -4 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke"; tail -4 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench_${TAG}_cfg2.log 2>&1; echo "== cfg2"; tail -1 gpurun_out/bench_${TAG}_cfg2.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.log 2>&1; echo "== reference arm"; tail -1 gpurun_out/bench_${TAG}_reference_arm.log | cut -c1-300
for c in ${BENCH_CONFIGS:-cfg3 cfg5 cfg4 cfg1}; do
  timeout 400 python bench.py --config $c --steps 50 --warmup 5 > gpurun_out/bench_${TAG}_$c.log 2>&1
  echo "== $c"; tail -1 gpurun_out/bench_${TAG}_$c.log | cut -c1-260
done
NCU_KERNELS="${NCU_KERNELS:-attn_fused_kernel attn_bwd_tc_kernel adam_kernel gemm_tc_kernel sce_reg_kernel ln_fwd_kernel ln_bwd_kernel}" PROFILE_TAG=$TAG bash tools/gpu_profile.sh > gpurun_out/profile_$TAG.log 2>&1
tail -30 gpurun_out/launches_summary_$TAG.txt
