#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "${TEST_FILTER:-sce or anchors or tiny or prefetch or trainer}" 2>&1 | tail -8 > gpurun_out/tests_sel.log
tail -4 gpurun_out/tests_sel.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_b.log 2>&1
echo "== bench"; tail -1 gpurun_out/bench_b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for r in d['rooflines'][:10]: print('  ', r['kernel'], r['launches_per_step'], 'alone', r['ms_alone'], r['bound'], r['frac'] and round(r['frac'],3))
"
