#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list of eager train steps, then --set full on the top kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG="${PROFILE_TAG:-r02}"
# launch list: 3 warm-up + 2 timed eager steps (about 150 launches each); steps 3 and 4 are summarised
[ "${SKIP_LAUNCH_LIST:-0}" = "1" ] || timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline \
    > gpurun_out/ncu_bench_stdout.log 2>&1
[ "${SKIP_LAUNCH_LIST:-0}" = "1" ] || { python tools/summarize_launches.py gpurun_out/launches_$TAG.csv 3 5 > gpurun_out/launches_summary_$TAG.txt 2>&1; tail -36 gpurun_out/launches_summary_$TAG.txt; }
for pat in ${NCU_KERNELS:-attn_fused_kernel attn_bwd_tc_kernel adam_kernel gemm_tc_persistent_kernel gemm_tc_kernel sce_kernel ln_fwd_kernel ln_bwd_kernel ln_bwd_reduce_kernel colsum_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-3} -f \
      -o gpurun_out/prof_${TAG}_$pat python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline \
      > gpurun_out/ncu_full_$pat.log 2>&1
  # summarise on the box and drop the report: gpurun copies back at most 64 MiB
  python tools/ncu_summary.py gpurun_out/prof_${TAG}_$pat.ncu-rep --traffic-json gpurun_out/${TAG}_traffic.json > gpurun_out/ncu_full_${TAG}_$pat.txt 2>&1
  rm -f gpurun_out/prof_${TAG}_$pat.ncu-rep
  head -16 gpurun_out/ncu_full_${TAG}_$pat.txt | cut -c1-150
done
