"""Warm per-kernel timing of vct_gemm shapes of the step (CUDA graph of N back-to-back launches, CUDA events)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from vct import lib as L
lib = L.load()
dev = "cuda"
SHAPES = [  # (tag, M, N, K, a_trans, b_trans, c_dtype)
    ("fwd  qkv   ", 1280, 2304, 768, 0, 0, L.BF16), ("fwd  proj  ", 1280, 768, 768, 0, 0, L.F32),
    ("fwd  ffn1  ", 1280, 2048, 768, 0, 0, L.BF16), ("fwd  ffn2  ", 1280, 768, 2048, 0, 0, L.F32),
    ("dgrad proj ", 1280, 768, 768, 0, 1, L.BF16), ("dgrad ffn2 ", 1280, 2048, 768, 0, 1, L.BF16),
    ("dgrad ffn1 ", 1280, 768, 2048, 0, 1, L.F32), ("dgrad qkv  ", 1280, 768, 2304, 0, 1, L.F32),
    ("wgrad proj ", 768, 768, 1280, 1, 1, L.F32), ("wgrad ffn1 ", 2048, 768, 1280, 1, 1, L.F32),
    ("wgrad ffn2 ", 768, 2048, 1280, 1, 1, L.F32), ("wgrad qkv  ", 2304, 768, 1280, 1, 1, L.F32),
    ("gen  fwd   ", 1280, 30522, 768, 0, 0, L.F32), ("gen  dgrad ", 1280, 768, 30522, 0, 1, L.F32),
    ("gen  wgrad ", 30522, 768, 1280, 1, 1, L.F32),
]
if os.environ.get("VARIANTS"):
    SHAPES = [("fwd N30720 ", 1280, 30720, 768, 0, 0, L.F32), ("fwd nobias ", 1280, 30522, 768, 0, 0, L.F32),
              ("fwd bf16out", 1280, 30522, 768, 0, 0, L.BF16), ("kk M30522  ", 30522, 768, 1280, 0, 0, L.F32),
              ("fwd K1280  ", 1280, 30522, 1280, 0, 0, L.F32), ("mn N30522  ", 1280, 30522, 768, 1, 1, L.F32),
              ("fwd M2560  ", 2560, 30522, 768, 0, 0, L.F32)]
reps = int(os.environ.get("REPS", "20"))
if os.environ.get("FIXED"):
    SHAPES = [("tiny 1 cta ", 128, 64, 64, 0, 0, L.F32), ("1cta K768  ", 128, 64, 768, 0, 0, L.F32), ("1cta K768 N256", 128, 256, 768, 0, 0, L.F32),
              ("120cta K64 ", 1280, 768, 64, 0, 0, L.F32), ("120cta K768", 1280, 768, 768, 0, 0, L.F32),
              ("8cta N256  ", 1024, 256, 768, 0, 0, L.F32), ("32cta N256 ", 4096, 256, 768, 0, 0, L.F32),
              ("90cta N256 ", 1280, 2304, 768, 0, 0, L.F32), ("90cta mfan ", 11520, 256, 768, 0, 0, L.F32),
              ("90cta nfan ", 128, 23040, 768, 0, 0, L.F32), ("148cta N256", 18944, 256, 768, 0, 0, L.F32),
              ("148 K3072  ", 18944, 256, 3072, 0, 0, L.F32)]
    # launch gap of a chain of trivial dependent kernels
    rng = torch.zeros(2, dtype=torch.int64, device=dev); hyper = torch.zeros(8, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        L.check(lib.vct_step_tick(rng.data_ptr(), hyper.data_ptr(), st.cuda_stream)); torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(100):
                L.check(lib.vct_step_tick(rng.data_ptr(), hyper.data_ptr(), st.cuda_stream))
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); graph.replay(); e1.record(st); torch.cuda.synchronize()
    print(f"chain of 100 trivial kernels in a graph: {e0.elapsed_time(e1)*10:.2f} us per kernel", flush=True)
for tag, M, N, K, at, bt, cd in SHAPES:
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
    g.bias = bias.data_ptr() if (not at and 'nobias' not in tag) else None
    g.impl = L.GEMM_TCGEN05
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            L.check(lib.vct_gemm(C.byref(g), st.cuda_stream))
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(reps):
                L.check(lib.vct_gemm(C.byref(g), st.cuda_stream))
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); graph.replay(); e1.record(st); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{tag} M{M:6d} N{N:6d} K{K:6d}  {us:8.2f} us  {2.0*M*N*K/us/1e6:8.1f} TFLOP/s", flush=True)
