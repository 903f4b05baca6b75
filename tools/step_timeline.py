"""Kernel timeline of one graph-replayed train step (CUPTI through torch.profiler): per-stream busy time, idle gaps
of the main lane and a per-kernel list in start order.  python tools/step_timeline.py [--batch 64] > gpurun_out/timeline.txt"""
import argparse, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200")); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--full", action="store_true")
args = ap.parse_args()
from model.MMT4Caption import MMT4Caption
from vct.synthetic import make_tokenizer_dir, shipped_model_config, synth_batch
from vct.trainer import CaptionTrainer
# under torchrun (WORLD_SIZE > 1) every rank trains data-parallel and rank 0 prints ITS timeline (NCCL kernels included)
rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
tok = os.path.join(ROOT, "gpurun_out", "_tok")
if rank == 0:
    make_tokenizer_dir(tok)
if world > 1:
    dist.barrier()
torch.manual_seed(666)
model = MMT4Caption(shipped_model_config(tok), device=dev).to(dev)
model.vct_precision, model.vct_gemm = "bf16", None
model.mode("caption"); model.train()
tr = CaptionTrainer(model, lr=1e-4, uniform_shapes=True)
x, vm, ids = synth_batch(args.batch, 12, 512, 21, seed=1234 + rank)
x, vm, ids = x.to(dev), vm.to(dev), ids.to(dev)
for _ in range(6):
    tr.step(x, vm, ids)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(x, vm, ids)
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank != 0:
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)
path = os.path.join(ROOT, "gpurun_out", "trace_step.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# split into steps at step_tick_kernel
starts = [i for i, e in enumerate(ev) if "step_tick" in e["name"]]
seg = ev[starts[1]:starts[2]] if len(starts) >= 3 else ev
t0 = seg[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in seg)
print(f"step: {len(seg)} kernels, {t1 - t0:.1f} us")
streams = collections.OrderedDict()
for e in seg:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in streams.items():
    busy = sum(e["dur"] for e in es)
    print(f"stream {s}: {len(es)} kernels, busy {busy:.1f} us, first {es[0]['ts']-t0:.1f}, last end {es[-1]['ts']+es[-1]['dur']-t0:.1f}")
main = max(streams.values(), key=len)
gap = 0.0; prev = None
for e in main:
    if prev is not None: gap += max(0.0, e["ts"] - prev)
    prev = e["ts"] + e["dur"]
print(f"main lane idle between kernels: {gap:.1f} us")
def short(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("vct::", "")
    return n[:58]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in seg:
    k = short(e["name"]); agg[k][0] += 1; agg[k][1] += e["dur"]
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:60s} x{c:3d} {d:8.1f} us")
sid = {s: i for i, s in enumerate(streams)}
for e in seg:
    print(f"{e['ts']-t0:9.1f} +{e['dur']:7.1f}  L{sid[e['args'].get('stream')]}  grid {str(e['args'].get('grid')):18s} {short(e['name'])}")
os.remove(path)
