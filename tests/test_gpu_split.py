"""The reference-precision tensor-core GEMM (csrc/gemm_split.cu: fp32 operands split into bf16 pieces, 3 / 6 cross terms
accumulated in fp32 by ONE tcgen05 launch) against float64 matmul on the same fp32 operands.

Tolerances (max |err| / sum_k |a||b|, printed per case): a product rebuilt from 6 terms is off by ~2^-25 |a||b| (the dropped
terms m*l, l*m, l*l); what remains is the fp32 accumulation of K' = 6K products in TMEM, i.e. the error class of any fp32
GEMM -- bound: 8x the error of torch's own fp32 (TF32 off) matmul on the same operands, + 1e-6.  3 terms drop
m*m ~ 2^-17 |a||b| -- bound 4e-5."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from vct import lib as L  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def lib():
    return L.load()


def stream():
    return torch.cuda.current_stream().cuda_stream


def split_gemm(lib, Ad, Bd, a_trans, b_trans, M, N, K, terms, bias=None):
    g = L.GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.a_dtype, g.lda, g.a_trans = Ad.data_ptr(), L.F32, Ad.stride(0), a_trans
    g.B, g.b_dtype, g.ldb, g.b_trans = Bd.data_ptr(), L.F32, Bd.stride(0), b_trans
    Cc = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
    g.C, g.c_dtype, g.ldc = Cc.data_ptr(), L.F32, N
    if bias is not None:
        g.bias = bias.data_ptr()
    g.impl = L.GEMM_TCGEN05_X3 if terms == 3 else L.GEMM_TCGEN05_X6
    need = int(lib.vct_gemm_split_workspace_bytes(M, N, K, a_trans, b_trans, terms))
    ws = torch.empty(need + 256, dtype=torch.uint8, device=DEV)
    g.split_ws, g.split_ws_bytes = ws.data_ptr(), ws.numel()
    sk = torch.empty(8 * M * ((N + 7) // 8 * 8), dtype=torch.float32, device=DEV)
    g.splitk_ws, g.splitk_ws_floats = sk.data_ptr(), sk.numel()
    L.check(lib.vct_gemm(C.byref(g), stream()), "vct_gemm(split)")
    torch.cuda.synchronize()
    return Cc


@pytest.mark.parametrize("terms", [3, 6])
@pytest.mark.parametrize("a_trans,b_trans", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 136, 768), (40, 216, 1003), (1280, 768, 2048)])
def test_split_gemm_matches_float64(lib, terms, a_trans, b_trans, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    # wide dynamic range: magnitudes over ~6 decades, so the low pieces matter
    A = torch.randn(M, K, generator=g) * torch.exp(3.0 * torch.randn(M, K, generator=g))
    B = torch.randn(N, K, generator=g) * torch.exp(3.0 * torch.randn(N, K, generator=g))
    Ad = (A.t().contiguous() if a_trans else A).to(DEV)
    Bd = (B.t().contiguous() if b_trans else B).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    got = split_gemm(lib, Ad, Bd, a_trans, b_trans, M, N, K, terms, bias=bias).double().cpu()
    want = A.double() @ B.double().t() + bias.double().cpu()
    scale = (A.double().abs() @ B.double().abs().t())            # sum_k |a||b|: what a relative product error scales with
    rel = float(((got - want).abs() / scale).max())
    torch.backends.cuda.matmul.allow_tf32 = False
    got32 = (A.to(DEV) @ B.to(DEV).t() + bias).double().cpu()
    rel32 = float(((got32 - want).abs() / scale).max())
    print(f"split GEMM x{terms} {M}x{N}x{K} trans=({a_trans},{b_trans}): max |err| / sum|a||b| = {rel:.3e} (torch fp32: {rel32:.3e})")
    assert rel <= (8.0 * rel32 + 1e-6 if terms == 6 else 4e-5), (rel, rel32)


@pytest.mark.parametrize("terms", [3, 6])
def test_split_pieces_reconstruct_the_operand(lib, terms):
    """vct_split_bf16 on its own: the pieces add back to x (to 2^-16 / 2^-24 |x|) and are laid out as documented."""
    rows, cols = 37, 100                                         # cols % 8 != 0: Kp = 104, zero padding
    x = (torch.randn(rows, cols) * torch.exp(2.0 * torch.randn(rows, cols))).to(DEV)
    Kp = 104
    dst = torch.full((rows, terms * Kp), float("nan"), dtype=torch.bfloat16, device=DEV)
    L.check(lib.vct_split_bf16(x.data_ptr(), cols, rows, cols, 1, 0, terms, dst.data_ptr(), terms * Kp, stream()), "vct_split_bf16")
    torch.cuda.synchronize()
    p = dst.float().view(rows, terms, Kp)
    assert float(p[:, :, cols:].abs().max()) == 0.0
    h, m = p[:, 0, :cols], p[:, 2, :cols]                        # A side: h h m (m h l)
    assert torch.equal(p[:, 1, :cols], h)
    assert torch.equal(h, x.to(torch.bfloat16).float())
    err2 = float(((h + m - x).abs() / x.abs().clamp_min(1e-30)).max())
    assert err2 <= 2.0 ** -15, err2
    if terms == 6:
        l = p[:, 5, :cols]
        err3 = float(((h.double() + m.double() + l.double() - x.double()).abs() / x.double().abs().clamp_min(1e-30)).max())
        assert err3 <= 2.0 ** -22, err3


def test_split_mode_train_step_matches_fp32_simt_engine(tokenizer_dir):
    """Same weights, same batch, dropout 0: the bf16x6 engine's loss and a deep gradient equal the SIMT fp32 engine's to
    fp32 round-off (both are fp32-class computations with different summation orders)."""
    from model.MMT4Caption import MMT4Caption
    from vct.synthetic import shipped_model_config, synth_batch
    out = {}
    for precision in ("fp32", "bf16x6", "bf16x3"):
        torch.manual_seed(666)
        m = MMT4Caption(shipped_model_config(tokenizer_dir, dropout=0.0), device=torch.device(DEV)).to(DEV)
        m.vct_precision = precision
        m.mode("caption")
        m.train()
        x, vm, tok = synth_batch(8, 12, 512, 21, padded=True)
        loss = m([x.to(DEV)], [vm.to(DEV)], tok.to(DEV))
        loss.backward()
        torch.cuda.synchronize()
        out[precision] = (float(loss), m.video_encoder.unify[0].weight.grad.detach().double().cpu(),
                          m.cap_decoder.generator.weight.grad.detach().double().cpu())
    l32, gu32, gg32 = out["fp32"]
    for precision, tl, tg in (("bf16x6", 2e-6, 2e-4), ("bf16x3", 2e-4, 2e-2)):
        l, gu, gg = out[precision]
        eu = float((gu - gu32).norm() / gu32.norm())
        eg = float((gg - gg32).norm() / gg32.norm())
        print(f"{precision}: loss {l:.7f} vs fp32 {l32:.7f}; grad rel-L2 unify {eu:.2e}, generator {eg:.2e}")
        assert abs(l - l32) <= tl * l32, (precision, l, l32)
        assert eu <= tg and eg <= tg, (precision, eu, eg)
