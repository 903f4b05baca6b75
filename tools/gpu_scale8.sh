cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
run() { # N tag extra-env
  N=$1; TAG=$2
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29566"
  env $3 timeout 240 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-rooflines > gpurun_out/scale_cfg2_n${N}_$TAG.log 2>&1
  echo "== cfg2 N=$N $TAG"; tail -1 gpurun_out/scale_cfg2_n${N}_$TAG.log | cut -c1-260
}
run 8 peer VCT_COMM=peer
run 8 nccl VCT_COMM=nccl
run 4 peer VCT_COMM=peer
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29567"
PRECISION=bf16 timeout 200 $TR tools/ddp_check.py > gpurun_out/ddp_check_peer_n8.log 2>&1; tail -2 gpurun_out/ddp_check_peer_n8.log | cut -c1-300
timeout 240 $TR bench.py --gpus 8 --config cfg4 --steps 100 --warmup 10 --no-rooflines > gpurun_out/scale_cfg4_n8_peer.log 2>&1; echo "== cfg4 N=8"; tail -1 gpurun_out/scale_cfg4_n8_peer.log | cut -c1-260
timeout 200 $TR tools/step_timeline.py > gpurun_out/timeline_n8_peer.txt 2>&1; grep -n "^step\|peer_" gpurun_out/timeline_n8_peer.txt | head -20 | cut -c1-130
