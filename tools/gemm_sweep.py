"""Tile sweep of the tcgen05 GEMM over every GEMM shape of the bench train step (B = 64: 1280 decoder rows,
832 encoder rows).  For each shape: the built-in choice ("auto") and every (block_n, ring, split-K) the kernel
supports, timed as a CUDA graph of REPS back-to-back launches (CUDA events, warm).  Writes
gpurun_out/gemm_sweep.json; the cost model in csrc/gemm_tc.cu is calibrated against it.

    python tools/gemm_sweep.py            # on the GPU box
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200"))
import torch  # noqa: E402
from vct import lib as L  # noqa: E402

lib = L.load()
dev = "cuda"
RD, RE, D, F, DIN = 1280, 832, 768, 2048, 512
SHAPES = []  # (tag, count per step, M, N, K, a_trans, b_trans, c_dtype)
for rows, n, who in ((RD, 3, "dec"), (RE, 1, "enc")):
    SHAPES += [
        (f"{who} fwd qkv", n, rows, 3 * D, D, 0, 0, L.BF16), (f"{who} fwd proj", n * (3 if who == "dec" else 1), rows, D, D, 0, 0, L.F32),
        (f"{who} fwd ffn1", n, rows, F, D, 0, 0, L.BF16), (f"{who} fwd ffn2", n, rows, D, F, 0, 0, L.F32),
        (f"{who} dgrad proj", n * (3 if who == "dec" else 1), rows, D, D, 0, 1, L.BF16), (f"{who} dgrad ffn2", n, rows, F, D, 0, 1, L.BF16),
        (f"{who} dgrad ffn1", n, rows, D, F, 0, 1, L.F32), (f"{who} dgrad qkv", n, rows, D, 3 * D, 0, 1, L.F32),
        (f"{who} wgrad proj", n * (3 if who == "dec" else 1), D, D, rows, 1, 1, L.F32), (f"{who} wgrad ffn1", n, F, D, rows, 1, 1, L.F32),
        (f"{who} wgrad ffn2", n, D, F, rows, 1, 1, L.F32), (f"{who} wgrad qkv", n, 3 * D, D, rows, 1, 1, L.F32),
    ]
SHAPES += [("enc fwd unify", 1, RE, D, DIN, 0, 0, L.F32), ("enc wgrad unify", 1, D, DIN, RE, 1, 1, L.F32),
           ("cross fwd kv", 3, RE, 2 * D, D, 0, 0, L.BF16), ("cross dgrad kv", 3, RE, D, 2 * D, 0, 1, L.F32),
           ("cross wgrad kv", 3, 2 * D, D, RE, 1, 1, L.F32),
           ("gen fwd", 1, RD, 30522, D, 0, 0, L.F32), ("gen dgrad", 1, RD, D, 30522, 0, 1, L.F32),
           ("gen wgrad", 1, 30522, D, RD, 1, 1, L.F32)]
REPS = int(os.environ.get("REPS", "12"))
only = os.environ.get("ONLY")
splitk = torch.empty(8 * 2304 * 2304, device=dev, dtype=torch.float32)


def time_cfg(g, st):
    for _ in range(2):
        L.check(lib.vct_gemm(C.byref(g), st.cuda_stream))
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=st):
        for _ in range(REPS):
            L.check(lib.vct_gemm(C.byref(g), st.cuda_stream))
    graph.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); graph.replay(); e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / REPS)
    return best


results = []
tot_auto = tot_best = 0.0
for tag, cnt, M, N, K, at, bt, cd in SHAPES:
    if only and only not in tag:
        continue
    lda = ((M + 7) // 8 * 8) if at else ((K + 7) // 8 * 8)
    ldb = ((N + 7) // 8 * 8) if bt else ((K + 7) // 8 * 8)
    A = torch.randn((K if at else M, lda), device=dev).to(torch.bfloat16)
    B = torch.randn((K if bt else N, ldb), device=dev).to(torch.bfloat16)
    ldc = (N + 7) // 8 * 8
    Cc = torch.empty((M, ldc), device=dev, dtype=torch.bfloat16 if cd == L.BF16 else torch.float32)
    bias = torch.randn(ldc, device=dev)
    g = L.GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.a_dtype, g.lda, g.a_trans = A.data_ptr(), L.BF16, lda, at
    g.B, g.b_dtype, g.ldb, g.b_trans = B.data_ptr(), L.BF16, ldb, bt
    g.C, g.c_dtype, g.ldc = Cc.data_ptr(), cd, ldc
    g.bias = bias.data_ptr() if not at else None
    g.impl = L.GEMM_TCGEN05
    g.splitk_ws, g.splitk_ws_floats = splitk.data_ptr(), splitk.numel()
    st = torch.cuda.Stream()
    row = {"tag": tag, "count": cnt, "M": M, "N": N, "K": K, "a_trans": at, "b_trans": bt, "cfgs": {}, "bad": []}
    Af = (A[:, :M].t() if at else A[:, :K]).float()
    Bf = (B[:, :N] if bt else B[:, :K].t()).float()
    ref = Af @ Bf
    if not at:
        ref += bias[:N]
    refn = float(ref.norm())
    with torch.cuda.stream(st):
        L.check(lib.vct_gemm_tune(0, 0, 0))
        row["auto_us"] = time_cfg(g, st)
        big = N > 8192 or M > 8192
        for bn in (64, 128, 256):
            for ring in ((0, 1) if bn <= 128 else (0, -1)):
                for sp in ((1,) if (big or ring == -1) else (1, 2, 3, 4, 6)):
                    if sp > 1 and (K + 63) // 64 < 4 * sp:
                        continue
                    if big and bn == 64:
                        continue
                    L.check(lib.vct_gemm_tune(bn, sp, ring))
                    Cc.zero_()
                    row["cfgs"][f"{bn}/{ring}/{sp}"] = time_cfg(g, st)
                    err = float((Cc[:, :N].float() - ref).norm()) / refn
                    if not err < 1e-2:
                        row["bad"].append((f"{bn}/{ring}/{sp}", err))
                        print(f"  !! {tag} {bn}/{ring}/{sp}: rel err {err:.3e}", flush=True)
                        row["cfgs"].pop(f"{bn}/{ring}/{sp}")
        L.check(lib.vct_gemm_tune(0, 0, 0))
    bk = min(row["cfgs"], key=row["cfgs"].get)
    row["best"], row["best_us"] = bk, row["cfgs"][bk]
    tot_auto += cnt * row["auto_us"]
    tot_best += cnt * row["best_us"]
    top = sorted(row["cfgs"].items(), key=lambda kv: kv[1])[:4]
    print(f"{tag:18s} x{cnt} M{M:6d} N{N:6d} K{K:6d} auto {row['auto_us']:7.2f} us | " +
          "  ".join(f"{k} {v:6.2f}" for k, v in top) + f" | {2.0*M*N*K/row['best_us']/1e6:6.0f} TF/s", flush=True)
    results.append(row)
print(f"per step: auto {tot_auto:.1f} us, best {tot_best:.1f} us")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gemm_sweep.json"), "w") as f:
    json.dump({"reps": REPS, "per_step_auto_us": tot_auto, "per_step_best_us": tot_best, "rows": results}, f, indent=1)
