"""Pipeline-phase timestamps (clock64) of CTA (0,0,0) of the tcgen05 tile kernel for a few shapes / tile configs."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200"))
import torch
from vct import lib as L
lib = L.load()
dev = "cuda"
trace = torch.zeros(128, dtype=torch.int64, device=dev)
splitk = torch.empty(8 * 2304 * 2304, device=dev, dtype=torch.float32)
CASES = [(1280, 768, 768, 0, 0, "64/0/1"), (1280, 768, 768, 0, 0, "128/0/1"), (1280, 768, 768, 0, 0, "256/0/1"),
         (1280, 768, 768, 0, 0, "64/1/1"), (1280, 2304, 768, 0, 0, "128/1/1"), (1280, 768, 2048, 0, 0, "64/0/1"),
         (1280, 768, 768, 0, 1, "64/0/1"), (768, 768, 1280, 1, 1, "64/0/1")]
ACT_CASES = {"gelu_fwd": (1280, 2048, 768, 0, 0, "128/1/1"), "gelu_bwd": (1280, 2048, 768, 0, 1, "128/1/1"), "plain": (1280, 2048, 768, 0, 0, "128/1/1")}
rng = torch.tensor([77, 5], dtype=torch.int64, device=dev)
for name, (M, N, K, at, bt, cfg) in ACT_CASES.items():
    lda = K; ldb = N if bt else K
    A = torch.randn((M, lda), device=dev).to(torch.bfloat16)
    B = torch.randn((K if bt else N, ldb), device=dev).to(torch.bfloat16)
    Cc = torch.empty((M, N), device=dev, dtype=torch.bfloat16); C2 = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    aux = torch.randn((M, N), device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    g = L.GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.a_dtype, g.lda, g.a_trans = A.data_ptr(), L.BF16, lda, at
    g.B, g.b_dtype, g.ldb, g.b_trans = B.data_ptr(), L.BF16, ldb, bt
    g.C, g.c_dtype, g.ldc = Cc.data_ptr(), L.BF16, N
    g.impl = L.GEMM_TCGEN05
    if name == "gelu_fwd":
        g.bias = bias.data_ptr(); g.act = L.ACT_GELU_FWD; g.C2, g.c2_dtype, g.ldc2 = C2.data_ptr(), L.BF16, N
        g.drop_p, g.rng_state, g.site = 0.3, rng.data_ptr(), 5
    elif name == "gelu_bwd":
        g.act = L.ACT_GELU_BWD; g.aux, g.aux_dtype, g.ld_aux = aux.data_ptr(), L.BF16, N
        g.drop_p, g.rng_state, g.site = 0.3, rng.data_ptr(), 5
    else:
        g.bias = bias.data_ptr()
    bn, ring, sp = map(int, cfg.split("/"))
    L.check(lib.vct_gemm_tune(bn, sp, ring)); L.check(lib.vct_gemm_trace(trace.data_ptr()))
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(6):
            L.check(lib.vct_gemm(C.byref(g), st.cuda_stream))
        torch.cuda.synchronize()
        t = trace.cpu().tolist(); z = t[0]
        L.check(lib.vct_gemm_trace(None))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(10):
                L.check(lib.vct_gemm(C.byref(g), st.cuda_stream))
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); graph.replay(); e1.record(st); torch.cuda.synchronize()
    print(f"{name:9s} {cfg}: {e0.elapsed_time(e1)*100:.2f} us/launch  pdl_wait {t[2]-z} last_commit {t[3]-z} acc_visible {t[4]-z} tmem_drained {t[5]-z} stores_issued {t[6]-z} exit {t[7]-z}", flush=True)
L.check(lib.vct_gemm_tune(0, 0, 0))
if os.environ.get("ACT_ONLY"):
    sys.exit(0)
for M, N, K, at, bt, cfg in CASES:
    lda = ((M + 7) // 8 * 8) if at else ((K + 7) // 8 * 8)
    ldb = ((N + 7) // 8 * 8) if bt else ((K + 7) // 8 * 8)
    A = torch.randn((K if at else M, lda), device=dev).to(torch.bfloat16)
    B = torch.randn((K if bt else N, ldb), device=dev).to(torch.bfloat16)
    ldc = (N + 7) // 8 * 8
    Cc = torch.empty((M, ldc), device=dev, dtype=torch.float32)
    bias = torch.randn(ldc, device=dev)
    g = L.GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.a_dtype, g.lda, g.a_trans = A.data_ptr(), L.BF16, lda, at
    g.B, g.b_dtype, g.ldb, g.b_trans = B.data_ptr(), L.BF16, ldb, bt
    g.C, g.c_dtype, g.ldc = Cc.data_ptr(), L.F32, ldc
    g.bias = bias.data_ptr() if not at else None
    g.impl = L.GEMM_TCGEN05
    g.splitk_ws, g.splitk_ws_floats = splitk.data_ptr(), splitk.numel()
    bn, ring, sp = map(int, cfg.split("/"))
    L.check(lib.vct_gemm_tune(bn, sp, ring))
    L.check(lib.vct_gemm_trace(trace.data_ptr()))
    st = torch.cuda.current_stream()
    for _ in range(6):      # back to back: the last launch is the warm, PDL-overlapped case
        L.check(lib.vct_gemm(C.byref(g), st.cuda_stream))
    torch.cuda.synchronize()
    t = trace.cpu().tolist()
    nkb = min(40, (K + 63) // 64 // sp)
    z = t[0]
    print(f"M{M} N{N} K{K} at{at} bt{bt} cfg {cfg}: setup {t[1]-z}  pdl_wait {t[2]-z}  last_commit {t[3]-z}  acc_visible {t[4]-z}  "
          f"tmem_drained {t[5]-z}  stores_issued {t[6]-z}  exit {t[7]-z}")
    if os.environ.get("VCT_LIB"):
        nkb = min(nkb, 20)
        print("   A issue   :", [t[16 + k] - z for k in range(nkb)])
        print("   B issue   :", [t[36 + k] - z for k in range(nkb)])
        print("   ops landed:", [t[56 + k] - z for k in range(nkb)])
        print("   committed :", [t[76 + k] - z for k in range(nkb)])
L.check(lib.vct_gemm_trace(None)); L.check(lib.vct_gemm_tune(0, 0, 0))
