#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29599"
PRECISION=bf16 timeout 200 $TR tools/ddp_check.py > gpurun_out/ddp_check_peer_n2.log 2>&1; tail -2 gpurun_out/ddp_check_peer_n2.log | cut -c1-300
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 --no-rooflines > gpurun_out/scale_cfg2_n2_final.log 2>&1; echo "== cfg2 N=2"; tail -1 gpurun_out/scale_cfg2_n2_final.log | cut -c1-260
timeout 200 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "gemm_grouped or peer" 2>&1 | tail -3
