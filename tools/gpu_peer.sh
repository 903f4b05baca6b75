cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${NGPU:-2}
timeout 300 python -m pytest tests/test_gpu_peer.py -q -x -p no:cacheprovider 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
PRECISION=bf16 timeout 200 $TR tools/ddp_check.py > gpurun_out/ddp_check_peer_n$N.log 2>&1; tail -2 gpurun_out/ddp_check_peer_n$N.log | cut -c1-400
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-rooflines > gpurun_out/scale_cfg2_n${N}_peer.log 2>&1; tail -1 gpurun_out/scale_cfg2_n${N}_peer.log | cut -c1-300
if [ "${WITH_CFG4:-0}" = "1" ]; then
timeout 300 $TR bench.py --gpus $N --config cfg4 --steps 100 --warmup 10 --no-rooflines > gpurun_out/scale_cfg4_n${N}_peer.log 2>&1; tail -1 gpurun_out/scale_cfg4_n${N}_peer.log | cut -c1-300
fi
timeout 200 $TR tools/step_timeline.py > gpurun_out/timeline_n${N}_peer.txt 2>&1; grep -n "^step\|peer_\|adam" gpurun_out/timeline_n${N}_peer.txt | head -30 | cut -c1-130
