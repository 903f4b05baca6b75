#!/bin/bash
# N-GPU runs (gpurun --gpus N): weak-scaling bench (cfg2), strong-scaling (cfg4), replica equivalence check.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "${RUN_TESTS:-0}" = "1" ]; then
  timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "${TEST_FILTER:-split or data}" -s 2>&1 | tail -25 > gpurun_out/tests_new.log; tail -12 gpurun_out/tests_new.log
fi
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-rooflines > gpurun_out/scale_cfg2_n$N.log 2>&1
echo "== cfg2 N=$N"; tail -1 gpurun_out/scale_cfg2_n$N.log | cut -c1-330
timeout 300 $TR bench.py --gpus $N --config cfg4 --steps 100 --warmup 10 --no-rooflines > gpurun_out/scale_cfg4_n$N.log 2>&1
echo "== cfg4 N=$N"; tail -1 gpurun_out/scale_cfg4_n$N.log | cut -c1-330
if [ "${RUN_DDP_CHECK:-1}" = "1" ]; then
  timeout 300 $TR tools/ddp_check.py > gpurun_out/ddp_check_n$N.log 2>&1; tail -4 gpurun_out/ddp_check_n$N.log
fi
