"""Greedy-decode throughput (BASELINE config 3: MSVD clip4clip model dims == MSR-VTT dims, batch 256, max_len 30,
random weights never emit [SEP] => 29 steps) through MMT4Caption.greedy_decode_ids (K/V-cached plan).
Prints captions/s for the fp32-exact path and the bf16 tensor-core path, and the token agreement between them."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200")); sys.path.insert(0, ROOT)
import torch
from model.MMT4Caption import MMT4Caption
from vct.synthetic import make_tokenizer_dir, shipped_model_config, synth_batch

dev = torch.device("cuda", 0)
tok = make_tokenizer_dir(os.path.join(ROOT, "gpurun_out", "_tok"))
B = int(os.environ.get("B", "256"))
x, vm, _ = synth_batch(B, 12, 512, 21, seed=1234)
xd, vd = x.to(dev), vm.to(dev)
out = {}
ys_ref = None
for precision in ("fp32", "bf16"):
    torch.manual_seed(666)
    m = MMT4Caption(shipped_model_config(tok), device=dev).to(dev)
    m.vct_precision = precision
    m.mode("caption"); m.eval()
    for sync_every in (1, 29):
        m.greedy_decode_ids([xd], [vd], max_len=30, sync_every=sync_every)          # warm (plans, kernels)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ys = m.greedy_decode_ids([xd], [vd], max_len=30, sync_every=sync_every)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        out[f"{precision}/sync_every={sync_every}"] = {"captions_per_s": B / dt, "ms": dt * 1e3, "tokens": int(ys.shape[1])}
    if ys_ref is None:
        ys_ref = ys.clone()
    else:
        out["bf16_vs_fp32_token_agreement"] = float((ys == ys_ref).float().mean())
        out["bf16_vs_fp32_first_token_agreement"] = float((ys[:, 1] == ys_ref[:, 1]).float().mean())
    del m
print(json.dumps({"workload": f"greedy decode B={B} max_len=30 (29 steps), shipped dims", **out}))
