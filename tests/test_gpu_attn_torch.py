"""The fused tcgen05 attention kernels (attn_fused.cu: forward of the three flavours and the backward) against plain
**fp32 PyTorch** arithmetic on the same (bf16-rounded) inputs -- not against the repo's own SIMT kernels (VERDICT r1,
"What's weak" 2).  The reference arithmetic is what torch's F.multi_head_attention_forward computes: packed
in-projection, scores * 1/sqrt(dh), additive -inf key-padding / causal masks, softmax, dropout on the probabilities,
P V.  Dropout is compared exactly: the keep mask of the kernels is a pure function of (seed, step, site, element index)
and is read back through the debug entry vct_dropout_mask (index space: probability row r = (b*H + h)*Lq + i owns
elements r*64 .. r*64+63).

Tolerances (bf16 operands, fp32 accumulation; q/k/v and the probabilities are rounded to bf16 before the second and
third contraction): saved projections |err| <= 2e-2 * (1 + |x|), probabilities |err| <= 4e-3, attention output and
gradients rel-L2 <= 1.2e-2 and |err| <= 4e-2 * max|ref|."""
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


def keep_mask(lib, rng, site, p, B, H, Lq, Lk):
    """[B, H, Lq, Lk] float keep multipliers (0 or 1/(1-p)) of the attention-probability dropout."""
    if p <= 0.0:
        return torch.ones(B, H, Lq, Lk, device=DEV)
    n = B * H * Lq * 64
    m = torch.zeros(n, dtype=torch.uint8, device=DEV)
    L.check(lib.vct_dropout_mask(m.data_ptr(), n, p, rng.data_ptr(), site, stream()))
    return m.view(B, H, Lq, 64)[..., :Lk].float() / (1.0 - p)


def torch_attention(q, k, v, key_pad, causal, keep):
    """q [B,Lq,H,dh], k/v [B,Lk,H,dh] fp32 -> (probs [B,H,Lq,Lk] pre-dropout, o [B,Lq,H,dh])."""
    B, Lq, H, dh = q.shape
    Lk = k.shape[1]
    s = torch.einsum("bihc,bjhc->bhij", q, k) / math.sqrt(dh)
    if key_pad is not None:
        s = s.masked_fill(key_pad.bool()[:, None, None, :], float("-inf"))
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Lq, Lk, dtype=torch.bool, device=q.device), 1), float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bhij,bjhc->bihc", p * keep, v)
    return p, o


def make_pad(B, Lk):
    kp = torch.zeros(B, Lk, dtype=torch.uint8)
    for b in range(1, B):
        if b % 5:
            kp[b, max(1, Lk - (b % 5)):] = 1          # position 0 ([CLS] / the global token) is never padded
    return kp.to(DEV)


def close(got, want, what, rel=1.2e-2, frac=4e-2):
    got, want = got.float(), want.float()
    assert torch.isfinite(got).all(), what
    e = float((got - want).norm() / (want.norm() + 1e-30))
    m = float((got - want).abs().max() / (want.abs().max() + 1e-30))
    assert e <= rel and m <= frac, (what, e, m)


SELF_CASES = [(64, 13, 768, 8, 0), (64, 20, 768, 8, 1), (7, 20, 768, 8, 1), (16, 33, 768, 8, 0), (128, 33, 768, 8, 0),
              (3, 64, 768, 8, 1), (5, 40, 768, 8, 1), (64, 13, 512, 8, 0), (6, 33, 512, 8, 1), (9, 1, 768, 8, 1)]


@pytest.mark.parametrize("B,Lq,d,H,causal", SELF_CASES)
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_fused_self_attention_vs_torch_fp32(lib, B, Lq, d, H, causal, p):
    dh = d // H
    g = torch.Generator().manual_seed(11 * B + Lq + d + causal)
    x = torch.randn(B * Lq, d, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(3 * d, d, generator=g) * 0.04).to(DEV, torch.bfloat16)
    b = (torch.randn(3 * d, generator=g) * 0.1).to(DEV)
    key_pad = make_pad(B, Lq)
    rng = torch.tensor([77, 5], dtype=torch.int64, device=DEV)
    site = 321
    m = L.MhaArgs()
    qkv = torch.full((B * Lq, 3 * d), float("nan"), device=DEV, dtype=torch.bfloat16)
    o = torch.full((B * Lq, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    probs = torch.zeros(B, H, Lq, Lq, device=DEV)
    m.B, m.L, m.Lk, m.d, m.H, m.dtype = B, Lq, Lq, d, H, L.BF16
    m.x, m.w_in, m.b_in = x.data_ptr(), w.data_ptr(), b.data_ptr()
    m.qkv, m.o, m.key_pad = qkv.data_ptr(), o.data_ptr(), key_pad.data_ptr()
    m.drop_p, m.rng_state, m.site = p, rng.data_ptr(), site
    m.probs = probs.data_ptr()
    m.gemm_impl = L.GEMM_TCGEN05
    fn = lib.vct_attn_dec_self_fwd if causal else lib.vct_attn_enc_self_fwd
    L.check(fn(C.byref(m), stream()))
    torch.cuda.synchronize()
    ref_qkv = x.float() @ w.float().t() + b
    torch.testing.assert_close(qkv.float(), ref_qkv, rtol=2e-2, atol=2e-2)
    # downstream reference from the bf16-rounded projections the kernel itself attends over
    qr, kr, vr = (t.reshape(B, Lq, H, dh) for t in qkv.float().split(d, dim=1))
    keep = keep_mask(lib, rng, site, p, B, H, Lq, Lq)
    pr, orf = torch_attention(qr, kr, vr, key_pad, causal, keep)
    torch.testing.assert_close(probs, pr, rtol=0, atol=4e-3)
    close(o.view(B, Lq, H, dh), orf, "attention output")


CROSS_CASES = [(64, 20, 13, 768, 8), (7, 20, 13, 768, 8), (16, 20, 33, 768, 8), (128, 20, 33, 768, 8), (5, 20, 13, 512, 8),
               (9, 25, 20, 768, 8), (64, 12, 13, 768, 8), (4, 40, 13, 768, 8), (4, 50, 40, 768, 8), (3, 64, 64, 512, 8)]


@pytest.mark.parametrize("B,Lq,Lk,d,H", CROSS_CASES)
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_fused_cross_attention_vs_torch_fp32(lib, B, Lq, Lk, d, H, p):
    dh = d // H
    g = torch.Generator().manual_seed(B + 3 * Lq + 5 * Lk + d)
    x = torch.randn(B * Lq, d, generator=g).to(DEV, torch.bfloat16)
    mem = torch.randn(B * Lk, d, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(3 * d, d, generator=g) * 0.04).to(DEV, torch.bfloat16)
    b = (torch.randn(3 * d, generator=g) * 0.1).to(DEV)
    kv = (mem.float() @ w[d:].float().t() + b[d:]).to(torch.bfloat16).contiguous()       # pre-projected, as the plans do
    rng = torch.tensor([78, 9], dtype=torch.int64, device=DEV)
    site = 654
    m = L.MhaArgs()
    q = torch.full((B * Lq, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    o = torch.full((B * Lq, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    probs = torch.zeros(B, H, Lq, Lk, device=DEV)
    m.B, m.L, m.Lk, m.d, m.H, m.dtype = B, Lq, Lk, d, H, L.BF16
    m.x, m.mem, m.w_in, m.b_in = x.data_ptr(), mem.data_ptr(), w.data_ptr(), b.data_ptr()
    m.qkv, m.kv, m.kv_ready, m.o = q.data_ptr(), kv.data_ptr(), 1, o.data_ptr()
    m.drop_p, m.rng_state, m.site = p, rng.data_ptr(), site
    m.probs = probs.data_ptr()
    m.gemm_impl = L.GEMM_TCGEN05
    L.check(lib.vct_attn_dec_cross_fwd(C.byref(m), stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(q.float(), x.float() @ w[:d].float().t() + b[:d], rtol=2e-2, atol=2e-2)
    qr = q.float().reshape(B, Lq, H, dh)
    kr, vr = (t.reshape(B, Lk, H, dh) for t in kv.float().split(d, dim=1))
    keep = keep_mask(lib, rng, site, p, B, H, Lq, Lk)
    pr, orf = torch_attention(qr, kr, vr, None, 0, keep)                 # cross-attention is never masked
    torch.testing.assert_close(probs, pr, rtol=0, atol=4e-3)
    close(o.view(B, Lq, H, dh), orf, "attention output")


BWD_CASES = [(64, 8, 20, 20, 96, 1), (64, 8, 13, 13, 96, 0), (64, 8, 20, 13, 96, 0), (7, 8, 20, 20, 96, 1), (16, 8, 20, 33, 96, 0),
             (16, 8, 33, 33, 96, 0), (128, 8, 33, 33, 96, 0), (5, 8, 20, 13, 64, 0), (9, 8, 25, 25, 64, 1), (3, 8, 64, 64, 96, 1),
             (4, 8, 40, 13, 96, 0), (4, 8, 50, 40, 64, 0), (6, 8, 33, 33, 64, 1)]


@pytest.mark.parametrize("B,H,Lq,Lk,dh,causal", BWD_CASES)
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_attention_backward_vs_torch_autograd_fp32(lib, B, H, Lq, Lk, dh, causal, p):
    g = torch.Generator().manual_seed(B + 7 * Lq + Lk + dh)
    d = H * dh
    q = torch.randn(B, Lq, H, dh, generator=g).to(DEV, torch.bfloat16)
    k = torch.randn(B, Lk, H, dh, generator=g).to(DEV, torch.bfloat16)
    v = torch.randn(B, Lk, H, dh, generator=g).to(DEV, torch.bfloat16)
    do = torch.randn(B, Lq, H, dh, generator=g).to(DEV, torch.bfloat16)
    key_pad = make_pad(B, Lk) if Lq == Lk else None
    rng = torch.tensor([91, 3], dtype=torch.int64, device=DEV)
    site = 17
    dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    dbias = torch.full((3 * d,), float("nan"), device=DEV)
    part = torch.empty(B, 3 * d, device=DEV)
    cnt = torch.zeros(H, dtype=torch.int32, device=DEV)
    a = L.AttnArgs()
    a.B, a.H, a.Lq, a.Lk, a.dh, a.dtype = B, H, Lq, Lk, dh, L.BF16
    a.q, a.q_ld, a.k, a.k_ld, a.v, a.v_ld = q.data_ptr(), d, k.data_ptr(), d, v.data_ptr(), d
    a.key_pad = key_pad.data_ptr() if key_pad is not None else None
    a.causal, a.scale = causal, 1.0 / math.sqrt(dh)
    a.drop_p, a.rng_state, a.site = p, rng.data_ptr(), site
    a.d_o, a.do_ld, a.dq, a.dq_ld, a.dk, a.dk_ld, a.dv, a.dv_ld = do.data_ptr(), d, dq.data_ptr(), d, dk.data_ptr(), d, dv.data_ptr(), d
    a.dbias, a.dbias_partials, a.dbias_counters = dbias.data_ptr(), part.data_ptr(), cnt.data_ptr()
    L.check(lib.vct_attn_bwd(C.byref(a), stream()), "bwd")
    torch.cuda.synchronize()
    assert int(cnt.abs().sum()) == 0
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    keep = keep_mask(lib, rng, site, p, B, H, Lq, Lk)
    _, o = torch_attention(qf, kf, vf, key_pad, causal, keep)
    o.backward(do.float())
    close(dq, qf.grad, "dq")
    close(dk, kf.grad, "dk")
    close(dv, vf.grad, "dv")
    ref_b = torch.cat([qf.grad.reshape(-1, d).sum(0), kf.grad.reshape(-1, d).sum(0), vf.grad.reshape(-1, d).sum(0)])
    assert float((dbias - ref_b).norm() / ref_b.norm()) < 1.5e-2


# ---- sequences longer than one tile (Lq or Lk > 64): the tiled kernels of csrc/attn_core.cu -------------------------
def keep_mask_long(lib, rng, site, p, B, H, Lq, Lk):
    """Long-path index space: probability row r owns ceil(Lk / 8) groups of 8 keys (include/vct.h, vct_attn_args)."""
    if p <= 0.0:
        return torch.ones(B, H, Lq, Lk, device=DEV)
    w = (Lk + 7) // 8 * 8
    n = B * H * Lq * w
    m = torch.zeros(n, dtype=torch.uint8, device=DEV)
    L.check(lib.vct_dropout_mask(m.data_ptr(), n, p, rng.data_ptr(), site, stream()))
    return m.view(B, H, Lq, w)[..., :Lk].float() / (1.0 - p)


LONG_CASES = [(2, 4, 70, 70, 96, 1, True), (2, 2, 100, 33, 64, 0, False), (1, 2, 20, 130, 96, 0, True),
              (2, 2, 65, 65, 32, 1, True), (1, 1, 300, 300, 128, 1, False), (3, 2, 1, 200, 96, 0, False),
              (1, 2, 257, 513, 64, 0, True), (2, 8, 81, 81, 96, 0, True)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,Lq,Lk,dh,causal,pad", LONG_CASES)
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_long_sequence_attention_vs_torch_fp32(lib, dtype, B, H, Lq, Lk, dh, causal, pad, p):
    """Forward (probabilities, output) and backward (dq, dk, dv) of the tiled kernels against fp32 PyTorch autograd on the
    same inputs, with the exact dropout mask.  fp32 storage: |err| <= 5e-5 + 2e-4 |ref|; bf16 storage: the tolerances of
    this file's header."""
    g = torch.Generator().manual_seed(3 * B + 7 * Lq + Lk + dh + causal)
    d = H * dh
    q = torch.randn(B, Lq, H, dh, generator=g).to(DEV, dtype)
    k = torch.randn(B, Lk, H, dh, generator=g).to(DEV, dtype)
    v = torch.randn(B, Lk, H, dh, generator=g).to(DEV, dtype)
    do = torch.randn(B, Lq, H, dh, generator=g).to(DEV, dtype)
    key_pad = None
    if pad:
        key_pad = torch.zeros(B, Lk, dtype=torch.uint8)
        key_pad[0, Lk - 5:] = 1                           # (position 0 is never padded)
        if B > 1:
            key_pad[1, Lk // 2:] = 1
        key_pad = key_pad.to(DEV)
    rng = torch.tensor([123, 4], dtype=torch.int64, device=DEV)
    site = 29
    o = torch.full_like(q, float("nan"))
    probs = torch.full((B, H, Lq, Lk), float("nan"), device=DEV)
    dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    stats = torch.empty(B * H * Lq, 4, device=DEV)
    a = L.AttnArgs()
    a.B, a.H, a.Lq, a.Lk, a.dh = B, H, Lq, Lk, dh
    a.dtype = L.BF16 if dtype == torch.bfloat16 else L.F32
    a.q, a.q_ld, a.k, a.k_ld, a.v, a.v_ld, a.o, a.o_ld = q.data_ptr(), d, k.data_ptr(), d, v.data_ptr(), d, o.data_ptr(), d
    a.key_pad = key_pad.data_ptr() if key_pad is not None else None
    a.causal, a.scale = causal, 1.0 / math.sqrt(dh)
    a.drop_p, a.rng_state, a.site = p, rng.data_ptr(), site
    a.probs = probs.data_ptr()
    a.d_o, a.do_ld, a.dq, a.dq_ld, a.dk, a.dk_ld, a.dv, a.dv_ld = do.data_ptr(), d, dq.data_ptr(), d, dk.data_ptr(), d, dv.data_ptr(), d
    a.row_stats = stats.data_ptr()
    L.check(lib.vct_attn_fwd(C.byref(a), stream()), "fwd")
    L.check(lib.vct_attn_bwd(C.byref(a), stream()), "bwd")
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    keep = keep_mask_long(lib, rng, site, p, B, H, Lq, Lk)
    if p > 0:
        rate = float((keep > 0).float().mean())
        assert abs(rate - (1 - p)) < 0.02, rate
    pr, orf = torch_attention(qf, kf, vf, key_pad, causal, keep)
    orf.backward(do.float())
    torch.testing.assert_close(probs, pr.detach(), rtol=1e-4, atol=1e-5)
    if dtype == torch.float32:
        for got, want, what in ((o, orf.detach(), "o"), (dq, qf.grad, "dq"), (dk, kf.grad, "dk"), (dv, vf.grad, "dv")):
            torch.testing.assert_close(got, want, rtol=2e-4, atol=5e-5, msg=lambda m, what=what: f"{what}: {m}")
    else:
        close(o, orf.detach(), "attention output")
        close(dq, qf.grad, "dq")
        close(dk, kf.grad, "dk")
        close(dv, vf.grad, "dv")


def test_long_sequence_backward_needs_its_workspace(lib):
    """Loud failure instead of a wild write: the tiled backward refuses to run without row_stats, and does not pretend to
    produce the in-projection bias gradient."""
    B, H, L_, dh = 1, 2, 80, 32
    d = H * dh
    t = [torch.zeros(B, L_, H, dh, device=DEV) for _ in range(7)]
    a = L.AttnArgs()
    a.B, a.H, a.Lq, a.Lk, a.dh, a.dtype = B, H, L_, L_, dh, L.F32
    a.q, a.q_ld, a.k, a.k_ld, a.v, a.v_ld = t[0].data_ptr(), d, t[1].data_ptr(), d, t[2].data_ptr(), d
    a.scale = 1.0
    a.d_o, a.do_ld, a.dq, a.dq_ld, a.dk, a.dk_ld, a.dv, a.dv_ld = t[3].data_ptr(), d, t[4].data_ptr(), d, t[5].data_ptr(), d, t[6].data_ptr(), d
    assert lib.vct_attn_bwd(C.byref(a), stream()) != 0
    assert b"row_stats" in lib.vct_last_error()
    a.Lq = a.Lk = 1025
    assert lib.vct_attn_fwd(C.byref(a), stream()) != 0
    assert b"1024" in lib.vct_last_error()
