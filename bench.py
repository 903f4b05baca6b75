#!/usr/bin/env python
"""Benchmark of the caption hot path (BASELINE.json metric: captions/sec of one train step; attention kernel
roofline fractions).

    python bench.py --gpus N --steps K --warmup W                    # our arm, BASELINE configs[1] (cfg2)
    python bench.py --config {cfg1,cfg1-literal,cfg3,cfg4,cfg5,cfg5-b128} ...   # the other BASELINE configs
    python bench.py --impl reference [--config ...] --steps K --warmup W         # the reference's own CPU path

Workloads (config.workload), all synthetic [B, T, 512] frame features + random token ids, S = 20 decoder positions:
  cfg2 (default)  shipped MSR-VTT clip4clip model (1 enc + 3 dec layers, d 768, 8 heads, FFN 2048, V 30522, dropout 0.3,
                  SCE loss), 64 captions per GPU, one train step = forward + backward + gradient exchange (N > 1) + Adam;
                  weak scaling
  cfg1 / cfg1-literal   B = 8 train step, shipped dims / the literal "2 enc + 2 dec, d 512" reading
  cfg3            greedy decode (predict_video path), B = 256, max_len 30 (29 steps), captions/s
  cfg4            train step, GLOBAL batch 512 over N GPUs (strong scaling: 512 / N per GPU)
  cfg5 / cfg5-b128      6 enc + 6 dec layers, d 768, T = 32 (M = 33): global batch 128 over N GPUs / 128 per GPU

Prints ONE JSON line (rank 0).  value = device-timed throughput with inputs resident in HBM; e2e = the same workload
driven through the public API with pinned HOST inputs (H2D copies + result read-back inside the timed region).
roofline = the kernel family with the LARGEST SHARE of the step (rooflines = every family), timed live with CUDA events;
traffic = dram bytes per launch from the committed ncu capture (profiles/r02_traffic.json) when present.
cpu_baseline / --impl reference = the reference's own modules (baseline/_ref, unmodified) on the host cores, train mode
with dropout 0.3 + torch.optim.Adam; gpu_reference = the same modules on this B200 through PyTorch's own sm_100 kernels
(fp32 with TF32 off, and bf16 autocast): the "existing Blackwell path".
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

DIN, S1, V = 512, 21, 30522
CONFIGS = {
    #            kind     enc dec d    T   batch rule
    "cfg1":         dict(kind="train", Le=1, Ld=3, d=768, T=12, per_gpu=8, scaling="weak",
                         name="msrvtt-clip4clip 1enc+3dec d768 h8 ff2048 V30522 dropout0.3, train step, B=8/GPU, T=12, S=20"),
    "cfg1-literal": dict(kind="train", Le=2, Ld=2, d=512, T=12, per_gpu=8, scaling="weak",
                         name="cfg1 literal reading: 2enc+2dec d512 h8 ff2048 V30522 dropout0.3, train step, B=8/GPU, T=12, S=20"),
    "cfg2":         dict(kind="train", Le=1, Ld=3, d=768, T=12, per_gpu=64, scaling="weak",
                         name="msrvtt-clip4clip 1enc+3dec d768 h8 ff2048 V30522 dropout0.3, train step, B=64/GPU, T=12, S=20"),
    "cfg3":         dict(kind="decode", Le=1, Ld=3, d=768, T=12, per_gpu=256, scaling="weak", max_len=30,
                         name="msvd-clip4clip 1enc+3dec d768 h8 ff2048 V30522, greedy decode, B=256/GPU, T=12, max_len=30 (29 steps)"),
    "cfg4":         dict(kind="train", Le=1, Ld=3, d=768, T=12, global_batch=512, scaling="strong",
                         name="msrvtt-clip4clip 1enc+3dec d768, train step, GLOBAL batch 512 (512/N per GPU), T=12, S=20"),
    "cfg5":         dict(kind="train", Le=6, Ld=6, d=768, T=32, global_batch=128, scaling="strong",
                         name="scaled 6enc+6dec d768 h8 ff2048, T=32 (M=33), train step, GLOBAL batch 128 (128/N per GPU), S=20"),
    "cfg5-b128":    dict(kind="train", Le=6, Ld=6, d=768, T=32, per_gpu=128, scaling="weak",
                         name="scaled 6enc+6dec d768 h8 ff2048, T=32 (M=33), train step, B=128/GPU, S=20"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "src": "fallback"}


# kernel family of the bench line -> kernel name in the ncu capture (profiles/r02_traffic.json, written by
# tools/ncu_summary.py --traffic-json from `ncu --set full` of `bench.py --steps 1 --warmup 3 --no-graph`)
FAMILY_KERNEL = {"vct_gemm:layers": "gemm_tc_kernel", "vct_gemm:generator": "gemm_tc_persistent_kernel", "vct_adam": "adam_kernel",
                 "vct_ln_residual_fwd": "ln_fwd_kernel", "vct_ln_residual_bwd": "ln_bwd_kernel", "vct_ln_bwd_reduce": "ln_bwd_reduce_kernel",
                 "vct_sce_typed": "sce_reg_kernel", "vct_colsum": "colsum_kernel", "vct_attn_bwd": "attn_bwd_tc_kernel",
                 "vct_attn_enc_self_fwd": "attn_fused_kernel", "vct_attn_dec_self_fwd": "attn_fused_kernel",
                 "vct_attn_dec_cross_fwd": "attn_fused_kernel"}


def ncu_traffic():
    """{kernel name: mean dram bytes (read + write) per launch} from the committed `ncu --set full` capture (cfg2 workload)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def tokenizer_dir():
    from vct.synthetic import make_tokenizer_dir
    return make_tokenizer_dir(os.path.join(ROOT, "gpurun_out", "_tok"))


def model_config(cfg, tokdir):
    from vct.synthetic import shipped_model_config
    return shipped_model_config(tokdir, embed_dim=cfg["d"], enc_layers=cfg["Le"], dec_layers=cfg["Ld"])


def per_gpu_batch(cfg, world, override=None):
    if override:
        return override
    if "per_gpu" in cfg:
        return cfg["per_gpu"]
    if cfg["global_batch"] % world:
        raise SystemExit(f"global batch {cfg['global_batch']} is not divisible by {world} ranks")
    return cfg["global_batch"] // world


# ---------------------------------------------------------------------------------------------------
# the reference on the host cores (reference arm / cpu_baseline) and on the GPU (gpu_reference)
# ---------------------------------------------------------------------------------------------------
def build_reference(cfg, device):
    """The UNMODIFIED reference MMT4Caption (baseline/_ref, or /root/reference in the build container) with the offline
    shims (tokenizer directory, clip stub).  Returns (model, kind) or (None, why)."""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        return None, "baseline/_ref not installed (python baseline/install_reference.py)"
    ref = ref_shims.import_reference_model()
    torch.manual_seed(666)
    model = ref.MMT4Caption.MMT4Caption(model_config(cfg, tokenizer_dir()), device=device)
    model.mode("caption")
    return model.to(device), "reference"


def reference_train_steps(model, x, vm, tok, steps, warmup, autocast=None):
    """Loop body of train.py:119-126 on pre-tokenised ids: forward (encoder + decoder + SCE loss), zero_grad, backward,
    Adam.  Returns the per-step wall times (s)."""
    model.train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-4, betas=(0.9, 0.999))
    cuda = x.is_cuda
    times = []
    for it in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with (torch.autocast("cuda", dtype=autocast) if autocast is not None else torch.autocast("cpu", enabled=False)):
            mem, _, _ = model.video_encoder([x], [vm])
            _, loss = model.cap_decoder(mem, tok, tok == 0)
        opt.zero_grad()
        loss.backward()
        opt.step()
        _ = loss.item()
        if cuda:
            torch.cuda.synchronize()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def reference_decode(model, x, vm, max_len, reps, warmup=1):
    """MMT4Caption.greedy_decode's loop (model/MMT4Caption.py:156-172) with ids kept as ids."""
    model.eval()
    times = []
    with torch.no_grad():
        for it in range(warmup + reps):
            if x.is_cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            memory, _, _ = model.video_encoder([x], [vm])
            ys = torch.full((x.shape[0], 1), 101, dtype=torch.long, device=x.device)
            for _ in range(max_len - 1):
                logits = model.cap_decoder.decode_word(memory, ys, None)
                ys = torch.cat([ys, logits.argmax(dim=1, keepdim=True)], dim=1)
                if bool((ys == 102).any(dim=1).all()):
                    break
            if x.is_cuda:
                torch.cuda.synchronize()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def cpu_baseline(cfg, steps, warmup, batch=None):
    """The reference's CPU path on a bounded sample of the workload.  Returns the cpu_baseline dict."""
    from vct.synthetic import synth_batch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cpu = torch.device("cpu")
    model, kind = build_reference(cfg, cpu)
    if model is None:
        return _cpu_baseline_port(cfg, steps, warmup, threads, kind)
    if cfg["kind"] == "decode":
        B = batch or 32
        x, vm, _ = synth_batch(B, cfg["T"], DIN, S1)
        t = reference_decode(model, x, vm, cfg["max_len"], max(1, min(steps, 3)), warmup=1)
        sec = statistics.median(t)
        return {"value": B / sec, "unit": "captions/s", "cores": threads, "kind": kind, "ms": sec * 1e3,
                "sample": f"{len(t)} greedy decodes of batch {B} (the reference's recompute-everything loop, max_len {cfg['max_len']}, "
                          f"fp32, eval), median"}
    B = batch or min(64, cfg.get("per_gpu", 64))
    x, vm, tok = synth_batch(B, cfg["T"], DIN, S1)
    t = reference_train_steps(model, x, vm, tok, steps, warmup)
    sec = statistics.median(t)
    return {"value": B / sec, "unit": "captions/s", "cores": threads, "kind": kind, "ms": sec * 1e3,
            "sample": f"{len(t)} train steps of batch {B} (the reference's own modules: fwd + bwd + torch.optim.Adam, fp32, "
                      f"train mode with dropout 0.3, pre-tokenised ids), median"}


def _cpu_baseline_port(cfg, steps, warmup, threads, why):
    """baseline/_ref missing: time the oracle restatement instead (kind "port")."""
    from oracle import vct_oracle as O
    from vct.synthetic import synth_batch
    from model.MMT4Caption import MMT4Caption
    torch.manual_seed(666)
    m = MMT4Caption(model_config(cfg, tokenizer_dir()), device=torch.device("cpu"))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    B = min(64, cfg.get("per_gpu", 64))
    x, vm, tok = synth_batch(B, cfg["T"], DIN, S1)
    state = {}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, grads = O.caption_grads(sd, x, vm, tok, 8, 8, 0.5)
        for k, gk in grads.items():
            mm, vv = state.get(k, (torch.zeros_like(gk), torch.zeros_like(gk)))
            sd[k], mm, vv = O.adam_step(sd[k], gk, mm, vv, it + 1, 1e-4)
            state[k] = (mm, vv)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return {"value": B / sec, "unit": "captions/s", "cores": threads, "kind": "port", "ms": sec * 1e3,
            "sample": f"{len(times)} train steps of batch {B} (oracle port: fwd+bwd+Adam, fp32, no dropout; {why}), median"}


def run_reference_arm(args, cfg):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cb = cpu_baseline(cfg, args.steps, args.warmup)
    metric = "captions/sec (greedy decode)" if cfg["kind"] == "decode" else "captions/sec (train step)"
    line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "captions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms"],
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_reference(cfg, B, dev, steps=12, warmup=4):
    """The reference's own modules on THIS GPU through PyTorch's stock sm_100 kernels (cuBLAS + SDPA + ATen): fp32 with
    TF32 off (the reference's precision) and under bf16 autocast.  Same synthetic batch, train mode, dropout 0.3, Adam."""
    from vct.synthetic import synth_batch
    out = {}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for tag, ac in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        try:
            model, kind = build_reference(cfg, dev)
            if model is None:
                return {"unavailable": kind}
            if cfg["kind"] == "decode":
                if ac is not None:
                    continue
                x, vm, _ = synth_batch(B, cfg["T"], DIN, S1)
                t = reference_decode(model, x.to(dev), vm.to(dev), cfg["max_len"], 3, warmup=1)
            else:
                x, vm, tok = synth_batch(B, cfg["T"], DIN, S1)
                t = reference_train_steps(model, x.to(dev), vm.to(dev), tok.to(dev), steps, warmup, autocast=ac)
            sec = statistics.median(t)
            out[tag] = {"value": B / sec, "unit": "captions/s", "ms": sec * 1e3, "batch": B}
        except Exception as e:                                  # the reference itself cannot run some modes
            out[tag] = {"error": f"{type(e).__name__}: {e}"[:200]}
        finally:
            model = None
            torch.cuda.empty_cache()
    out["how"] = ("unmodified reference modules (baseline/_ref) on this GPU, PyTorch %s stock kernels, wall clock around "
                  "synchronised steps, median of %d" % (torch.__version__, steps))
    return out


# ---------------------------------------------------------------------------------------------------
# per-kernel timing of our arm
# ---------------------------------------------------------------------------------------------------
def profile_calls(plans):
    """Per-launch device time of every C-ABI call of the given plans (CUDA events on the launching stream, eager).
    Returns [(name, ms, args)] in issue order."""
    stream = torch.cuda.current_stream()
    out = []
    # keep the GPU busy while the host queues every launch + event, so the event deltas are device time
    torch.cuda._sleep(int(4e7))
    for plan in plans:
        for name, fn, a, _lane in plan.calls:
            if name == "join" or name.startswith("py:"):
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = fn(*a, stream.cuda_stream)
            e1.record(stream)
            if rc != 0:
                raise RuntimeError(f"{name} failed rc={rc}")
            out.append((name, e0, e1, a))
    torch.cuda.synchronize()
    return [(n, e0.elapsed_time(e1), a) for n, e0, e1, a in out]


def lib_fn(name):
    from vct import lib as L
    return getattr(L.load(), name.split(":")[0])


def graph_time_ms(name, a, reps=10):
    """One kernel alone: a CUDA graph of `reps` back-to-back launches of the call, CUDA events around the replay (the
    eager per-launch figure also contains the event / launch gap of ~3 us)."""
    fn = lib_fn(name)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(2):
            fn(*a, st.cuda_stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn(*a, st.cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
    return best


def family_of(name):
    base = name.split(":")[0]
    if base == "vct_gemm":
        return "vct_gemm:generator" if "generator" in name else "vct_gemm:layers"
    if base == "vct_gemm_grouped":
        return "vct_gemm:layers"
    if base == "vct_attn_bwd":
        return "vct_attn_bwd"
    return base


def algorithmic_work(name, a, eng):
    """(bytes, flops) of one call -- SURVEY section 8d formulas, bf16 = 2 bytes, fp32 = 4 (stated per kernel in DESIGN.md)."""
    base = name.split(":")[0]
    es = 2.0 if eng.precision == "bf16" else 4.0
    if base == "vct_gemm":
        g = a[0]._obj
        fl = 2.0 * g.M * g.N * g.K
        cb = 4.0 if g.c_dtype == 0 else 2.0
        by = es * (g.M * g.K + g.N * g.K) + cb * g.M * g.N + (2.0 * g.M * g.N if g.C2 else 0.0)
        return by, fl
    if base == "vct_gemm_grouped":                          # a[0] = ctypes array of vct_gemm_args, a[1] = count
        by = fl = 0.0
        for i in range(a[1]):
            g = a[0][i]
            fl += 2.0 * g.M * g.N * g.K
            by += es * (g.M * g.K + g.N * g.K) + 4.0 * g.M * g.N
        return by, fl
    if base in ("vct_attn_enc_self_fwd", "vct_attn_dec_self_fwd"):
        o = a[0]._obj
        B, L, d = o.B, o.L, o.d
        return es * (B * L * d + 3 * d * d + 3 * d + B * L * d + 3 * B * L * d) + B * L, B * (6.0 * L * d * d + 4.0 * L * L * d)
    if base == "vct_attn_dec_cross_fwd":
        o = a[0]._obj
        B, S, M, d = o.B, o.L, o.Lk, o.d
        return es * (B * S * d + B * M * 2 * d + d * d + d + 2 * B * S * d), B * (2.0 * S * d * d + 4.0 * S * M * d)
    if base in ("vct_attn_bwd", "vct_attn_fwd"):
        o = a[0]._obj
        B, Lq, Lk, d = o.B, o.Lq, o.Lk, o.H * o.dh
        if base == "vct_attn_fwd":
            return es * B * d * (2 * Lq + 2 * Lk), 4.0 * B * Lq * Lk * d
        return es * B * d * (3 * Lq + 4 * Lk), 10.0 * B * Lq * Lk * d
    if base == "vct_ln_residual_fwd":
        R, d = a[10], a[11]
        # reads r (+ x when there is a residual), writes s (when kept), y fp32 and the compute-dtype copy
        n_in = 2 if a[0] else 1
        return R * d * (4.0 * n_in + (4.0 if a[7] else 0.0) + 4.0 + (es if a[5] else 0.0)), 0.0
    if base == "vct_ln_residual_bwd":
        R, d = a[13], a[14]
        return R * d * (4.0 + 4.0 + 4.0 + (es if a[6] else 0.0)), 0.0
    if base == "vct_sce_typed":
        B, S, Vv = a[5], a[6], a[7]
        lb = 4.0 if a[1] == 0 else 2.0                      # logits storage type
        return B * S * Vv * (lb + (es if a[14] else 0.0)), 0.0
    if base == "vct_adam":
        n = a[5]
        return n * (16.0 + 12.0 + (2.0 if a[4] else 0.0)), 0.0
    return 0.0, 0.0


def attn_key(name, a):
    base = name.split(":")[0]
    if base == "vct_attn_bwd":
        o = a[0]._obj
        return "vct_attn_bwd:" + ("cross" if "cross" in name else ("enc_self" if "enc" in name else "dec_self"))
    return base


def kernel_rooflines(eng, plans, peaks, adam_ms, extra_adam=True):
    """Every kernel family of the step: eager per-launch CUDA-event time (share of the step), each distinct call also
    timed ALONE in a CUDA graph (achieved rate against the measured peaks)."""
    profile_calls(plans)                               # warm
    reps = [profile_calls(plans) for _ in range(3)]
    names = [r[0] for r in reps[0]]
    calls = [r[2] for r in reps[0]]
    med = [statistics.median(rep[i][1] for rep in reps) for i in range(len(names))]
    total = sum(med) + (adam_ms or 0.0)
    traffic = ncu_traffic()
    fam = {}
    alone_cache = {}
    for n, m, a in zip(names, med, calls):
        f = family_of(n)
        by, fl = algorithmic_work(n, a, eng)
        if f.startswith(("vct_gemm", "vct_attn", "vct_ln", "vct_sce", "vct_colsum")):
            if n not in alone_cache:
                alone_cache[n] = graph_time_ms(n, a)
            alone = alone_cache[n]
        else:
            alone = m
        e = fam.setdefault(f, {"launches": 0, "eager_ms": 0.0, "alone_ms": 0.0, "bytes": 0.0, "flops": 0.0})
        e["launches"] += 1
        e["eager_ms"] += m
        e["alone_ms"] += alone
        e["bytes"] += by
        e["flops"] += fl
    if adam_ms:
        a = eng.arena
        fam["vct_adam"] = {"launches": 1, "eager_ms": adam_ms, "alone_ms": adam_ms,
                           "bytes": a.numel * (16.0 + 12.0 + (2.0 if eng.precision == "bf16" else 0.0)), "flops": 0.0}
    out = []
    for f, e in fam.items():
        t = e["alone_ms"] * 1e-3
        t_hbm, t_tc = e["bytes"] / (peaks["hbm_gbs"] * 1e9), e["flops"] / (peaks["tflops"] * 1e12)
        bound = "tensor" if t_tc > t_hbm else "hbm"
        ach = (e["flops"] / t / 1e12) if bound == "tensor" else (e["bytes"] / t / 1e9)
        peak = peaks["tflops"] if bound == "tensor" else peaks["hbm_gbs"]
        # dram bytes per LAUNCH (mean over the captured launches of the family's kernel), next to the algorithmic bytes per
        # launch; only meaningful for the workload the capture was taken on (cfg2)
        tr = traffic.get(FAMILY_KERNEL.get(f, ""))
        out.append({"kernel": f, "launches_per_step": e["launches"], "share_of_step": e["eager_ms"] / total,
                    "ms_eager": round(e["eager_ms"], 4), "ms_alone": round(e["alone_ms"], 4), "bound": bound,
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                    "frac": ach / peak if t > 0 and (e["bytes"] or e["flops"]) else None,
                    "bytes": e["bytes"], "flops": e["flops"], "traffic": tr,
                    "bytes_per_launch": e["bytes"] / max(1, e["launches"]), "peak_source": peaks["src"]})
    out.sort(key=lambda r: -r["share_of_step"])
    # attention flavours, one line each (BASELINE metric "attn kernel HBM GB/s vs peak")
    attn = {}
    for n, m, a in zip(names, med, calls):
        if not n.startswith("vct_attn"):
            continue
        k = attn_key(n, a)
        by, fl = algorithmic_work(n, a, eng)
        r = attn.setdefault(k, {"kernel": k, "launches_per_step": 0, "ms": alone_cache[n], "ms_eager_with_gap": m,
                                "bytes": by, "flops": fl})
        r["launches_per_step"] += 1
    attn_out = []
    for k, r in attn.items():
        t = r["ms"] * 1e-3
        t_hbm, t_tc = r["bytes"] / (peaks["hbm_gbs"] * 1e9), r["flops"] / (peaks["tflops"] * 1e12)
        r.update({"achieved_gbs": r["bytes"] / t / 1e9, "achieved_tflops": r["flops"] / t / 1e12,
                  "ideal_us_hbm": t_hbm * 1e6, "ideal_us_tensor": t_tc * 1e6, "bound": "hbm" if t_hbm >= t_tc else "tensor",
                  "frac": max(t_hbm, t_tc) / t})
        r["ms"], r["ms_eager_with_gap"] = round(r["ms"], 5), round(r["ms_eager_with_gap"], 5)
        attn_out.append(r)
    breakdown = {r["kernel"]: r["ms_eager"] for r in out}
    return out, attn_out, breakdown, len(names)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)     # 200 x ~1.5 ms: long enough for several nvidia-smi clock samples
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="captions per GPU (overrides the config's rule)")
    ap.add_argument("--precision", default=os.environ.get("VCT_PRECISION", "bf16"))
    ap.add_argument("--gemm", default=os.environ.get("VCT_GEMM"))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-rooflines", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference_arm(args, cfg)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from model.MMT4Caption import MMT4Caption
    from vct.synthetic import synth_batch

    if rank == 0:
        tokenizer_dir()
    if world > 1:
        dist.barrier()
    tokdir = tokenizer_dir()
    torch.manual_seed(666)
    model = MMT4Caption(model_config(cfg, tokdir), device=dev).to(dev)
    model.vct_precision, model.vct_gemm = args.precision, args.gemm
    model.mode("caption")
    B = per_gpu_batch(cfg, world, args.batch)
    T = cfg["T"]
    x, vm, tok = synth_batch(B, T, DIN, S1, seed=1234 + rank)
    xd, vd, td = x.to(dev), vm.to(dev), tok.to(dev)
    xh, vh, th = x.pin_memory(), vm.pin_memory(), tok.pin_memory()
    verbose = os.environ.get("VCT_BENCH_VERBOSE") == "1"

    def note(msg):
        if verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    peaks = measured_peaks()
    decode = cfg["kind"] == "decode"
    if decode:
        model.eval()
        eng = model._engine()
        max_len = cfg["max_len"]

        def step_dev():
            return model.greedy_decode_ids([xd], [vd], max_len=max_len, sync_every=max_len - 1)

        def step_host():
            ys = model.greedy_decode_ids([xh], [vh], max_len=max_len, sync_every=max_len - 1)
            return ys.cpu()
        d2h = B * max_len * 8
        h2d = x.numel() * 4 + vm.numel()
        trainer = None
    else:
        from vct.trainer import CaptionTrainer
        model.train()
        trainer = CaptionTrainer(model, lr=1e-4, betas=(0.9, 0.999), use_graph=not args.no_graph, uniform_shapes=True)
        eng = trainer.engine

        def step_dev():
            return trainer.step(xd, vd, td)

        def step_host():
            # the public API as a training loop uses it: this step consumes the batch whose H2D copy was started during the
            # previous step, then the copy of the next batch is started and the loss is read back (host sync every step)
            loss = trainer.step(xh, vh, th)
            trainer.prefetch(xh, vh, th)
            return loss.item()
        d2h = 4
        h2d = x.numel() * 4 + vm.numel() + tok.numel() * 8

    # ---- warm-up (also builds plans / captures the graphs: needs >= 3 steps) -------------------------
    note("model built")
    for i in range(max(args.warmup, 3)):
        step_dev()
        torch.cuda.synchronize()
    sync_all()
    note("warm-up done")
    # ---- value: device-resident inputs ----------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        last = step_dev()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = eng.launches - launches0
    final = float(last.item()) if not decode else int(last.shape[1])
    note("timed region done")
    # ---- e2e: pinned host inputs, H2D + result read-back every step -----------------------------------
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    note("e2e done")
    clocks = sampler.stop() if rank == 0 else None
    comm = "single GPU"
    if world > 1 and trainer is not None:
        comm = "nccl"
        if getattr(trainer, "peer", None) is not None:
            comm = "peer-memory kernels (NVLink, vct_peer_allreduce_bf16 / vct_peer_allgather), one CUDA graph per step"
            if trainer.peer.status() != 0:                       # a cross-GPU barrier timed out: the numbers are garbage
                raise SystemExit(f"[bench rank {rank}] peer-memory exchange reported a barrier time-out")

    # ---- rooflines (rank 0): per-launch CUDA-event times of one eager pass + every distinct call alone in a graph ------
    rooflines, attn_roof, breakdown, phase_ms, roofline = None, None, None, None, None
    if rank == 0 and not args.no_rooflines:
        if decode:
            dws = eng.decode_workspace(B, T, max_len)
            plans = [eng.plan_decode_step(dws, max_len - 2)]
            rooflines, attn_roof, breakdown, _ = kernel_rooflines(eng, plans, peaks, None)
        else:
            ws = eng.workspace(B, T, S1 - 1, True)
            plans = [eng.plan_forward(ws, fused_grad=True, part="all"), eng.plan_backward(ws, sce_first=False, part="all")]
            e_ad0, e_ad1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.adam(1.0); torch.cuda.synchronize()
            e_ad0.record(); eng.adam(1.0); e_ad1.record(); torch.cuda.synchronize()
            rooflines, attn_roof, breakdown, _ = kernel_rooflines(eng, plans, peaks, e_ad0.elapsed_time(e_ad1))
            fwd = plans[0]
            bwd = eng.plan_backward(ws, sce_first=False, part="all", fuse_adam=(world == 1 and trainer.fuse_adam))
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            acc = [0.0, 0.0]
            for rep in range(4):
                torch.cuda._sleep(int(2e7))
                ev[0].record(); eng.run(fwd); ev[1].record(); eng.run(bwd); ev[2].record()
                torch.cuda.synchronize()
                if rep:
                    acc[0] += ev[0].elapsed_time(ev[1]) / 3
                    acc[1] += ev[1].elapsed_time(ev[2]) / 3
            phase_ms = {"forward": round(acc[0], 4), "backward_incl_side_lanes": round(acc[1], 4),
                        "n_forward_calls": len(fwd), "n_backward_calls": len(bwd)}
        if args.config != "cfg2" or B != 64 or eng.precision != "bf16":
            for r in rooflines:                             # the ncu capture is of the cfg2 workload
                r["traffic"] = None
        top = rooflines[0]                                  # the family with the largest share of the step
        roofline = {k: top[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "share_of_step",
                                        "launches_per_step", "ms_alone", "peak_source")}
    # ---- reference arms (rank 0, N = 1 only) -------------------------------------------------------------
    cpu, gref = None, None
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(cfg, 4 if not decode else 1, 1)
            cpu.pop("ms", None)
        if not args.no_gpu_reference:
            model_ref_B = B
            del model
            torch.cuda.empty_cache()
            gref = gpu_reference(cfg, model_ref_B, dev)
    if rank == 0:
        ms_step = ms_total / args.steps
        metric = "captions/sec (greedy decode)" if decode else "captions/sec (train step)"
        line = {"metric": metric, "value": B * world / (ms_step * 1e-3), "unit": "captions/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
                "dtype": "bf16" if eng.precision == "bf16" else ("f32" if eng.precision == "fp32" else eng.precision),
                "data": "synthetic",
                "config": {"workload": cfg["name"], "name": args.config, "global_batch": B * world, "per_gpu_batch": B,
                           "parallelism": f"dp{world}", "gemm": {0: "simt", 1: "tcgen05", 2: "tcgen05 split-bf16 x3", 3: "tcgen05 split-bf16 x6"}[eng.gemm_impl],
                           "cuda_graph": not args.no_graph, "gradient_exchange": comm,
                           "l2": "no explicit flush: one step streams ~2.5 GB of weights/moments/activations, 20x the 126 MB L2"},
                "e2e": {"value": B * world * args.steps / float(e2e_s.item()), "unit": "captions/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, ("tokens" if decode else "loss"): final, "clocks": clocks, "roofline": roofline,
                "rooflines": rooflines, "attn_rooflines": attn_roof, "cpu_baseline": cpu, "gpu_reference": gref,
                "kernel_ms": breakdown, "phase_ms": phase_ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
