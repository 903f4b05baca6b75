#!/bin/bash
# 2-GPU box: in-process peer-collective tests, peer-path bench sanity, the reference's own train.py through the drop-in,
# and the ncu capture of the cfg 5 cross-attention kernel (BASELINE configs[4]).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py -q -x -p no:cacheprovider 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588"
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 --no-rooflines > gpurun_out/scale_cfg2_n2_final.log 2>&1; echo "== cfg2 N=2"; tail -1 gpurun_out/scale_cfg2_n2_final.log | cut -c1-300
NGPU=2 bash tools/gpu_reference_train.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused_kernel -s 40 -c 14 -f -o gpurun_out/prof_cfg5_attn \
    python bench.py --config cfg5 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference --no-rooflines > gpurun_out/ncu_cfg5_attn.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_cfg5_attn.ncu-rep > gpurun_out/ncu_full_cfg5_attn_fused_kernel.txt 2>&1
rm -f gpurun_out/prof_cfg5_attn.ncu-rep
grep -A6 "Kernel Name" gpurun_out/ncu_full_cfg5_attn_fused_kernel.txt | grep "Kernel Name\|duration\|Grid" | cut -c1-170 | head -30
