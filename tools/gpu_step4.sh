#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-gpu-reference > gpurun_out/bench.log 2>&1
echo "== bench"; tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for r in d['rooflines'][:9]: print('  ', r['kernel'], r['launches_per_step'], 'alone', r['ms_alone'], r['bound'], r['frac'] and round(r['frac'],3), 'traffic', r['traffic'])
"
VCT_LOGITS=fp32 timeout 300 python bench.py --steps 100 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-rooflines > gpurun_out/bench_fp32logits.log 2>&1
echo "== bench fp32 logits"; tail -1 gpurun_out/bench_fp32logits.log | cut -c1-200
