#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list of eager train steps, then --set full on the top kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# 3 warm-up steps are skipped (~175 launches each + shadow cast); 2 steps captured
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 360 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline \
    > gpurun_out/ncu_bench_stdout.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1
tail -32 gpurun_out/launches_summary.txt
for pat in ${NCU_KERNELS:-adam_kernel gemm_tc_persistent_kernel gemm_tc_kernel attn_fwd_kernel attn_bwd_kernel ln_fwd_kernel ln_bwd_kernel sce_kernel colsum_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 8 -c 2 -f \
      -o gpurun_out/prof_$pat python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline \
      > gpurun_out/ncu_full_$pat.log 2>&1
  ls -la gpurun_out/prof_$pat.ncu-rep 2>&1 | awk '{print $5, $9}'
done
VCT_FUSED_ATTN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 4 -c 2 -f \
    -o gpurun_out/prof_attn_fused python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_attn_fused.log 2>&1
ls -la gpurun_out/prof_attn_fused.ncu-rep | awk '{print $5, $9}'
