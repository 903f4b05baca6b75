"""Print-only probe of the tcgen05 GEMM (all majors x a few shapes): max error vs fp64, so one GPU call
shows which layout combinations are right."""
import ctypes as C, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from vct import lib as L
import test_gpu_kernels as T
lib = L.load()
for (M, N, K) in [(128, 64, 64), (128, 256, 64), (128, 128, 256), (256, 512, 768), (200, 136, 72), (1280, 768, 30528)]:
    for at, bt in [(0, 0), (0, 1), (1, 0), (1, 1)]:
        Ad, Bd, Ar, Br = T.make_operands(M, N, K, at, bt, torch.bfloat16)
        try:
            Cc, _ = T.run_gemm(lib, Ad, Bd, at, bt, M, N, K, impl=L.GEMM_TCGEN05)
            want = Ar.double() @ Br.double().t()
            err = (Cc.double() - want).abs().max().item()
            nan = int(torch.isnan(Cc).sum().item())
            print(f"M{M} N{N} K{K} at{at} bt{bt}: max|err| {err:.3e} (scale {want.abs().max().item():.1f}) nan {nan}", flush=True)
        except Exception as e:
            print(f"M{M} N{N} K{K} at{at} bt{bt}: EXC {e}", flush=True)
            break
