#!/bin/bash
# Round-2 measurement sweep on one B200 (through gpurun): every BASELINE config through bench.py, then the ncu passes.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG="${PROFILE_TAG:-r02}"
for c in ${BENCH_CONFIGS:-cfg1 cfg1-literal cfg3 cfg4 cfg5 cfg5-b128}; do
  timeout 600 python bench.py --config $c --steps ${BENCH_STEPS:-50} --warmup 5 > gpurun_out/bench_${TAG}_$c.log 2>&1
  echo "== $c"; tail -1 gpurun_out/bench_${TAG}_$c.log | cut -c1-400
done
[ "${SKIP_PROFILE:-0}" = "1" ] || PROFILE_TAG=$TAG bash tools/gpu_profile.sh
