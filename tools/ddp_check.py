"""N-GPU vs 1-GPU equivalence of the native trainer (run under torchrun): every rank trains on its shard of a
global batch for 3 steps (per-slice NCCL all-reduce + Adam overlapped with backward); rank 0 then trains a second
replica alone on the FULL batch and compares parameters.  Un-padded inputs and dropout 0, so the mean of the
per-rank gradients equals the global-batch gradient."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from model.MMT4Caption import MMT4Caption
from vct.synthetic import make_tokenizer_dir, shipped_model_config, synth_batch
from vct.trainer import CaptionTrainer

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tok = os.path.join(ROOT, "gpurun_out", "_tok")
if rank == 0:
    make_tokenizer_dir(tok)
dist.barrier()
precision = os.environ.get("PRECISION", "fp32")
B = 8 * world
x, vm, ids = synth_batch(B, 12, 512, 21, seed=99, padded=False)


def build():
    torch.manual_seed(666)
    m = MMT4Caption(shipped_model_config(tok, dropout=0.0), device=dev).to(dev)
    m.vct_precision = precision
    m.mode("caption"); m.train()
    return m


m = build()
tr = CaptionTrainer(m, lr=1e-3, use_graph=True, uniform_shapes=os.environ.get("UNIFORM", "1") == "1")
sl = slice(rank * 8, (rank + 1) * 8)
losses = []
for _ in range(4):
    losses.append(tr.step(x[sl].to(dev), vm[sl].to(dev), ids[sl].to(dev)).item())
torch.cuda.synchronize()
mean_loss = torch.tensor(losses, device=dev)
dist.all_reduce(mean_loss)
mean_loss /= world
# replicas must be bit-identical
flat = tr.engine.arena.p32
ref = flat.clone()
dist.broadcast(ref, 0)
same = bool(torch.equal(ref, flat))
ok = torch.tensor([1 if same else 0], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    m1 = build()
    t1 = CaptionTrainer(m1, lr=1e-3, use_graph=True, world_size=1)
    l1 = [t1.step(x.to(dev), vm.to(dev), ids.to(dev)).item() for _ in range(4)]
    torch.cuda.synchronize()
    a, b = tr.engine.arena.p32, t1.engine.arena.p32
    rel = float((a - b).norm() / b.norm())
    maxabs = float((a - b).abs().max())
    print(f"ddp_check world={world} precision={precision}: replicas identical={bool(ok.item())}; "
          f"loss N-GPU {mean_loss.tolist()} vs 1-GPU {l1}; params rel-L2 {rel:.3e} max|diff| {maxabs:.3e}", flush=True)
    assert bool(ok.item())
    assert rel < (1e-5 if precision == "fp32" else 2e-3), rel
    print("DDP_CHECK_OK", flush=True)
dist.barrier()
dist.destroy_process_group()
