"""Run one vct_gemm shape with forced tile configs (for ncu captures):  python tools/gemm_one.py M N K at bt cfg[,cfg...]  (cfg = bn/ring/splits)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200"))
import torch
from vct import lib as L
lib = L.load()
M, N, K, at, bt = map(int, sys.argv[1:6])
cfgs = sys.argv[6].split(",")
dev = "cuda"
lda = ((M + 7) // 8 * 8) if at else ((K + 7) // 8 * 8)
ldb = ((N + 7) // 8 * 8) if bt else ((K + 7) // 8 * 8)
A = torch.randn((K if at else M, lda), device=dev).to(torch.bfloat16)
B = torch.randn((K if bt else N, ldb), device=dev).to(torch.bfloat16)
ldc = (N + 7) // 8 * 8
Cc = torch.empty((M, ldc), device=dev, dtype=torch.float32)
bias = torch.randn(ldc, device=dev)
splitk = torch.empty(8 * 2304 * 2304, device=dev, dtype=torch.float32)
g = L.GemmArgs()
g.M, g.N, g.K = M, N, K
g.A, g.a_dtype, g.lda, g.a_trans = A.data_ptr(), L.BF16, lda, at
g.B, g.b_dtype, g.ldb, g.b_trans = B.data_ptr(), L.BF16, ldb, bt
g.C, g.c_dtype, g.ldc = Cc.data_ptr(), L.F32, ldc
g.bias = bias.data_ptr() if not at else None
g.impl = L.GEMM_TCGEN05
g.splitk_ws, g.splitk_ws_floats = splitk.data_ptr(), splitk.numel()
st = torch.cuda.current_stream()
for cfg in cfgs:
    bn, ring, sp = map(int, cfg.split("/"))
    L.check(lib.vct_gemm_tune(bn, sp, ring))
    for _ in range(3):
        L.check(lib.vct_gemm(C.byref(g), st.cuda_stream))
    torch.cuda.synchronize()
print("done")
