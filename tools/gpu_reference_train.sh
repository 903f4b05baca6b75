#!/bin/bash
# The reference's OWN train.py (unmodified copy under baseline/_ref) driven through the drop-in `model` package on N GPUs
# under torchrun (--multi_gpu: DistributedDataParallel + DistributedSampler + loss all-reduce + rank-0 validation + greedy
# sample), on a small synthetic MSR-VTT-layout dataset.  VERDICT r1 "What's missing" #7.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
OUT=$PWD/gpurun_out/synth
rm -rf $OUT
python tools/make_synth_dataset.py $OUT $PWD/baseline/_ref > gpurun_out/synth_cfg.txt 2>&1 || { cat gpurun_out/synth_cfg.txt; exit 1; }
VCT_RUN_DIR=$PWD/gpurun_out/run timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
    tools/run_reference.py $PWD/baseline/_ref train.py -c $OUT/config.json --multi_gpu -ws $N > gpurun_out/reference_train_n$N.log 2>&1
echo "rc=$?"
grep -v "Warning\|warn\|^\s*$" gpurun_out/reference_train_n$N.log | tail -25 | cut -c1-220
rm -rf $OUT/checkpoint $OUT/feats gpurun_out/run    # keep gpurun_out under the 64 MiB copy-back limit (the reference saves a 300 MB checkpoint)
