#!/bin/bash
# Round 2, long-sequence attention + LayerNorm load hoisting: the new tests, a sanitizer pass over the tiled kernels, a short bench.
#   gpurun --timeout 420 -- 'bash tools/gpu_long.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
P="python -m pytest -q -p no:cacheprovider -m gpu"
timeout 400 $P tests/test_gpu_attn_torch.py tests/test_gpu_parity.py tests/test_gpu_data.py tests/test_gpu_kernels.py \
    -k "long or ln_ or tiny_forward" 2>&1 | tail -120 > gpurun_out/long_tests.log
timeout 200 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_long.log 2>&1
SEL='long_sequence and dtype0 and 0.3-2-4-70-70'
timeout 110 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_attn_torch.py -k "$SEL" > gpurun_out/long_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/long_racecheck.log
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_attn_torch.py -k "$SEL" > gpurun_out/long_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/long_memcheck.log
echo "== tests"; tail -25 gpurun_out/long_tests.log
echo "== racecheck"; tail -4 gpurun_out/long_racecheck.log
echo "== memcheck"; tail -4 gpurun_out/long_memcheck.log
echo "== bench"; tail -1 gpurun_out/bench_long.log | cut -c1-300
