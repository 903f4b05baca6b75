#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list of one eager train step, then --set full on the top kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G="${VCT_GEMM:-tcgen05}"
# 3 warm-up steps (~170 launches each incl. vct_cast) are skipped; 2 steps captured
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 340 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --gemm $G \
    > gpurun_out/ncu_bench_stdout.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1
tail -45 gpurun_out/launches_summary.txt
for pat in ${NCU_KERNELS:-gemm_tc_kernel attn_bwd_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s 12 -c 3 -f \
      -o gpurun_out/prof_$pat python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --gemm $G \
      > gpurun_out/ncu_full_$pat.log 2>&1
  ls -la gpurun_out/prof_$pat.ncu-rep
done
