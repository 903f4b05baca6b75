"""Kernel-level parity: every C-ABI entry point against fp32 PyTorch arithmetic / the oracle on the same
seeded inputs (run on the B200 box: ``pytest -m gpu``).  Tolerances are written next to each check."""
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


def tdt(code):
    return torch.bfloat16 if code == L.BF16 else torch.float32


def run_gemm(lib, A, B, a_trans, b_trans, M, N, K, c_dtype=L.F32, impl=L.GEMM_SIMT, **kw):
    g = L.GemmArgs()
    g.M, g.N, g.K = M, N, K
    dt = L.BF16 if A.dtype == torch.bfloat16 else L.F32
    g.A, g.a_dtype, g.lda, g.a_trans = A.data_ptr(), dt, A.stride(0), a_trans
    g.B, g.b_dtype, g.ldb, g.b_trans = B.data_ptr(), dt, B.stride(0), b_trans
    ldc = kw.get("ldc", N)
    Cc = torch.full((M, ldc), float("nan"), dtype=tdt(c_dtype), device=DEV)
    g.C, g.c_dtype, g.ldc = Cc.data_ptr(), c_dtype, ldc
    C2 = None
    if kw.get("c2_dtype") is not None:
        C2 = torch.full((M, N), float("nan"), dtype=tdt(kw["c2_dtype"]), device=DEV)
        g.C2, g.c2_dtype, g.ldc2 = C2.data_ptr(), kw["c2_dtype"], N
    keep = []
    for name in ("bias", "row_table", "addend", "aux", "rng_state"):
        t = kw.get(name)
        if t is not None:
            keep.append(t)
            setattr(g, name, t.data_ptr())
    g.row_period = kw.get("row_period", 0)
    g.ld_addend = N
    g.ld_aux = N
    g.aux_dtype = L.BF16 if (kw.get("aux") is not None and kw["aux"].dtype == torch.bfloat16) else L.F32
    g.act = kw.get("act", L.ACT_NONE)
    g.drop_p = kw.get("drop_p", 0.0)
    g.site = kw.get("site", 0)
    g.impl = impl
    L.check(lib.vct_gemm(C.byref(g), stream()), "vct_gemm")
    torch.cuda.synchronize()
    return Cc, C2


def make_operands(M, N, K, a_trans, b_trans, dtype, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    Ad = (A.t().contiguous() if a_trans else A).to(DEV, dtype)
    Bd = (B.t().contiguous() if b_trans else B).to(DEV, dtype)
    # reference operands after the storage rounding
    Ar = (Ad.t() if a_trans else Ad).float()
    Br = (Bd.t() if b_trans else Bd).float()
    return Ad, Bd, Ar, Br


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("a_trans,b_trans", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 136, 72), (37, 211, 40), (260, 48, 1000)])
def test_gemm_simt_layouts(lib, dtype, a_trans, b_trans, M, N, K):
    if dtype == torch.bfloat16 and (M % 8 or N % 8 or K % 8):
        M, N, K = (M + 7) // 8 * 8, (N + 7) // 8 * 8, (K + 7) // 8 * 8
    Ad, Bd, Ar, Br = make_operands(M, N, K, a_trans, b_trans, dtype)
    Cc, _ = run_gemm(lib, Ad, Bd, a_trans, b_trans, M, N, K)
    want = Ar.double() @ Br.double().t()
    # fp32 accumulation of exactly-represented products: error ~ K * eps * |terms|
    torch.testing.assert_close(Cc.double(), want, rtol=1e-4, atol=1e-4 * math.sqrt(K))


@pytest.mark.parametrize("shapes", [[(768, 2048, 1280), (2048, 768, 1280), (768, 768, 1280), (1536, 768, 832), (2304, 768, 1280)],
                                    [(200, 136, 72)], [(768, 512, 832), (136, 264, 64), (128, 256, 200)]])
def test_gemm_grouped_matches_individual_gemms(lib, shapes):
    """vct_gemm_grouped (one persistent launch over the tiles of up to 8 weight-gradient GEMMs, bf16 operands stored
    [K, M] / [K, N]) against fp32 matmul of the same bf16 operands, and a group the fast path does not cover."""
    args = (L.GemmArgs * len(shapes))()
    keep, want = [], []
    for i, (M, N, K) in enumerate(shapes):
        Ad, Bd, Ar, Br = make_operands(M, N, K, 1, 1, torch.bfloat16, seed=10 + i)
        Cc = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
        g = args[i]
        g.M, g.N, g.K = M, N, K
        g.A, g.a_dtype, g.lda, g.a_trans = Ad.data_ptr(), L.BF16, Ad.stride(0), 1
        g.B, g.b_dtype, g.ldb, g.b_trans = Bd.data_ptr(), L.BF16, Bd.stride(0), 1
        g.C, g.c_dtype, g.ldc = Cc.data_ptr(), L.F32, N
        g.impl = L.GEMM_TCGEN05
        keep.append((Ad, Bd, Cc))
        want.append(Ar.double() @ Br.double().t())
    for _ in range(2):
        L.check(lib.vct_gemm_grouped(args, len(shapes), stream()), "vct_gemm_grouped")
    torch.cuda.synchronize()
    for (_, _, Cc), w, (M, N, K) in zip(keep, want, shapes):
        torch.testing.assert_close(Cc.double(), w, rtol=1e-4, atol=1e-4 * math.sqrt(K))
    # a group with a bias (not a weight-gradient shape) falls back to one vct_gemm per problem
    bias = torch.randn(shapes[0][1], device=DEV)
    args[0].bias = bias.data_ptr()
    keep[0][2].fill_(float("nan"))
    L.check(lib.vct_gemm_grouped(args, len(shapes), stream()), "vct_gemm_grouped (fallback)")
    torch.cuda.synchronize()
    torch.testing.assert_close(keep[0][2].double(), want[0] + bias.double(), rtol=1e-4, atol=1e-4 * math.sqrt(shapes[0][2]))


def test_gemm_epilogue_bias_table_addend_and_bf16_out(lib):
    M, N, K = 130, 96, 64
    Ad, Bd, Ar, Br = make_operands(M, N, K, 0, 0, torch.float32, seed=1)
    g = torch.Generator().manual_seed(2)
    bias = torch.randn(N, generator=g).to(DEV)
    table = torch.randn(13, N, generator=g).to(DEV)
    addend = torch.randn(M, N, generator=g).to(DEV)
    Cc, C2 = run_gemm(lib, Ad, Bd, 0, 0, M, N, K, bias=bias, row_table=table, row_period=13, addend=addend,
                      c2_dtype=L.BF16)
    rows = torch.arange(M, device=DEV) % 13
    want = Ar @ Br.t() + bias + table[rows] + addend
    torch.testing.assert_close(Cc, want, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(C2.float(), want, rtol=1e-2, atol=1e-2)     # bf16 rounding of the same values


def test_gemm_gelu_forward_and_backward_epilogues(lib):
    M, N, K = 64, 128, 48
    Ad, Bd, Ar, Br = make_operands(M, N, K, 0, 0, torch.float32, seed=3)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(4)).to(DEV)
    z, h = run_gemm(lib, Ad, Bd, 0, 0, M, N, K, bias=bias, act=L.ACT_GELU_FWD, c2_dtype=L.F32)
    zw = Ar @ Br.t() + bias
    torch.testing.assert_close(z, zw, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(h, torch.nn.functional.gelu(zw), rtol=1e-4, atol=1e-4)
    # backward epilogue: C = acc * gelu'(aux)
    zz = zw.clone().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    Cc, _ = run_gemm(lib, Ad, Bd, 0, 0, M, N, K, act=L.ACT_GELU_BWD, aux=zw.contiguous())
    torch.testing.assert_close(Cc, (Ar @ Br.t()) * zz.grad, rtol=1e-4, atol=1e-4)


def test_gemm_dropout_mask_is_consistent_between_forward_backward_and_debug_entry(lib):
    M, N, K = 64, 256, 32
    Ad, Bd, Ar, Br = make_operands(M, N, K, 0, 0, torch.float32, seed=5)
    rng = torch.tensor([1234, 7], dtype=torch.int64, device=DEV)
    p = 0.3
    z, h = run_gemm(lib, Ad, Bd, 0, 0, M, N, K, act=L.ACT_GELU_FWD, c2_dtype=L.F32, drop_p=p, rng_state=rng, site=42)
    mask = torch.empty(M * N, dtype=torch.uint8, device=DEV)
    L.check(lib.vct_dropout_mask(mask.data_ptr(), M * N, p, rng.data_ptr(), 42, stream()))
    torch.cuda.synchronize()
    keep = mask.view(M, N).bool()
    want = torch.nn.functional.gelu(z) * keep / (1 - p)
    torch.testing.assert_close(h, want, rtol=1e-5, atol=1e-6)
    rate = keep.float().mean().item()
    assert abs(rate - (1 - p)) < 0.02, rate                     # keep-rate statistics
    Cc, _ = run_gemm(lib, Ad, Bd, 0, 0, M, N, K, act=L.ACT_GELU_BWD, aux=z.contiguous(), drop_p=p, rng_state=rng, site=42)
    zz = z.clone().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    torch.testing.assert_close(Cc, (Ar @ Br.t()) * zz.grad * keep / (1 - p), rtol=1e-4, atol=1e-4)
    # a different step or site gives a different mask
    rng2 = torch.tensor([1234, 8], dtype=torch.int64, device=DEV)
    mask2 = torch.empty(M * N, dtype=torch.uint8, device=DEV)
    L.check(lib.vct_dropout_mask(mask2.data_ptr(), M * N, p, rng2.data_ptr(), 42, stream()))
    torch.cuda.synchronize()
    assert (mask2 != mask).float().mean().item() > 0.2


def torch_attention(q, k, v, key_pad, causal, scale):
    s = torch.einsum("bihc,bjhc->bhij", q, k) * scale
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :].bool(), float("-inf"))
    if causal:
        L_ = s.shape[-1]
        s = s + torch.triu(torch.full((L_, L_), float("-inf"), device=s.device), diagonal=1)
    p = torch.softmax(s, dim=-1)
    return torch.einsum("bhij,bjhc->bihc", p, v), p


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,Lq,Lk,dh,causal,pad", [(3, 8, 13, 13, 96, 0, True), (4, 8, 20, 20, 96, 1, True),
                                                     (2, 8, 20, 13, 64, 0, False), (2, 4, 33, 33, 96, 0, True),
                                                     (2, 2, 7, 40, 16, 0, True), (5, 2, 1, 9, 24, 0, False)])
def test_attention_forward_backward(lib, dtype, B, H, Lq, Lk, dh, causal, pad):
    g = torch.Generator().manual_seed(B * 100 + Lq)
    d = H * dh
    q = torch.randn(B, Lq, H, dh, generator=g).to(DEV, dtype)
    k = torch.randn(B, Lk, H, dh, generator=g).to(DEV, dtype)
    v = torch.randn(B, Lk, H, dh, generator=g).to(DEV, dtype)
    do = torch.randn(B, Lq, H, dh, generator=g).to(DEV, dtype)
    key_pad = None
    if pad:
        key_pad = torch.zeros(B, Lk, dtype=torch.uint8)
        for b in range(1, B):
            key_pad[b, Lk - b:] = 1                      # column 0 never padded
        key_pad = key_pad.to(DEV)
    o = torch.full_like(q, float("nan"))
    probs = torch.zeros(B, H, Lq, Lk, device=DEV)
    dq, dk, dv = torch.full_like(q, float("nan")), torch.full_like(k, float("nan")), torch.full_like(v, float("nan"))
    a = L.AttnArgs()
    a.B, a.H, a.Lq, a.Lk, a.dh = B, H, Lq, Lk, dh
    a.dtype = L.BF16 if dtype == torch.bfloat16 else L.F32
    a.q, a.q_ld, a.k, a.k_ld, a.v, a.v_ld, a.o, a.o_ld = q.data_ptr(), d, k.data_ptr(), d, v.data_ptr(), d, o.data_ptr(), d
    a.key_pad = key_pad.data_ptr() if key_pad is not None else None
    a.causal, a.scale = causal, 1.0 / math.sqrt(dh)
    a.probs = probs.data_ptr()
    a.d_o, a.do_ld, a.dq, a.dq_ld, a.dk, a.dk_ld, a.dv, a.dv_ld = do.data_ptr(), d, dq.data_ptr(), d, dk.data_ptr(), d, dv.data_ptr(), d
    L.check(lib.vct_attn_fwd(C.byref(a), stream()), "fwd")
    L.check(lib.vct_attn_bwd(C.byref(a), stream()), "bwd")
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ow, pw = torch_attention(qf, kf, vf, key_pad, causal, a.scale)
    (ow * do.float()).sum().backward()
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(probs, pw.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(o.float(), ow.detach(), **tol)
    torch.testing.assert_close(dq.float(), qf.grad, **tol)
    torch.testing.assert_close(dk.float(), kf.grad, **tol)
    torch.testing.assert_close(dv.float(), vf.grad, **tol)


@pytest.mark.parametrize("R,d", [(40, 768), (1280, 768), (33, 512), (7, 32), (100, 48), (9, 1024)])
def test_ln_residual_forward_backward(lib, R, d):
    g = torch.Generator().manual_seed(R + d)
    x = torch.randn(R, d, generator=g).to(DEV)
    r = torch.randn(R, d, generator=g).to(DEV)
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(d, generator=g)).to(DEV)
    dy = torch.randn(R, d, generator=g).to(DEV)
    y = torch.empty(R, d, device=DEV)
    y_c = torch.empty(R, d, device=DEV, dtype=torch.bfloat16)
    s = torch.empty(R, d, device=DEV)
    mean, rstd = torch.empty(R, device=DEV), torch.empty(R, device=DEV)
    L.check(lib.vct_ln_residual_fwd(x.data_ptr(), r.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                    y_c.data_ptr(), L.BF16, s.data_ptr(), mean.data_ptr(), rstd.data_ptr(), R, d, 0.0,
                                    None, 0, stream()))
    xs = (x + r).requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yw = torch.nn.functional.layer_norm(xs, (d,), gp, bp, 1e-5)
    yw.backward(dy)
    torch.cuda.synchronize()
    torch.testing.assert_close(s, xs.detach(), rtol=0, atol=0)
    torch.testing.assert_close(y, yw.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(y_c.float(), yw.detach(), rtol=1e-2, atol=1e-2)
    ds = torch.empty(R, d, device=DEV)
    dr = torch.empty(R, d, device=DEV)
    dgamma, dbeta, dbias = torch.empty(d, device=DEV), torch.empty(d, device=DEV), torch.empty(d, device=DEV)
    nws = lib.vct_ln_bwd_workspace_floats(R, d)
    partials = torch.empty(nws, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    for _ in range(2):   # twice: the self-resetting counter must allow re-use
        L.check(lib.vct_ln_residual_bwd(dy.data_ptr(), s.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                        ds.data_ptr(), dr.data_ptr(), L.F32, dgamma.data_ptr(), dbeta.data_ptr(),
                                        dbias.data_ptr(), partials.data_ptr(), counter.data_ptr(), R, d, 0.0, None, 0, stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(ds, xs.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dr, xs.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dgamma, gp.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dbeta, bp.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dbias, xs.grad.sum(0), rtol=1e-4, atol=1e-4)
    assert int(counter.item()) == 0


def test_ln_without_residual_and_with_dropout(lib):
    R, d, p = 64, 768, 0.3
    g = torch.Generator().manual_seed(9)
    x = torch.randn(R, d, generator=g).to(DEV)
    r = torch.randn(R, d, generator=g).to(DEV)
    gamma, beta = torch.ones(d, device=DEV), torch.zeros(d, device=DEV)
    y = torch.empty(R, d, device=DEV)
    L.check(lib.vct_ln_residual_fwd(None, r.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), None, L.F32, None,
                                    None, None, R, d, p, None, 0, stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(y, torch.nn.functional.layer_norm(r, (d,)), rtol=1e-5, atol=1e-5)
    rng = torch.tensor([5, 3], dtype=torch.int64, device=DEV)
    s = torch.empty(R, d, device=DEV)
    L.check(lib.vct_ln_residual_fwd(x.data_ptr(), r.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), None, L.F32,
                                    s.data_ptr(), None, None, R, d, p, rng.data_ptr(), 77, stream()))
    mask = torch.empty(R * d, dtype=torch.uint8, device=DEV)
    L.check(lib.vct_dropout_mask(mask.data_ptr(), R * d, p, rng.data_ptr(), 77, stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(s, x + r * mask.view(R, d) / (1 - p), rtol=1e-6, atol=1e-6)


def test_embedding_forward_backward(lib):
    B, S, d, V = 5, 9, 64, 300
    g = torch.Generator().manual_seed(11)
    ids = torch.randint(0, V, (B, S + 1), generator=g)
    ids[0, 3] = 0
    ids[1, 2] = ids[1, 5]                       # duplicate token: scatter-add
    E = torch.randn(V, d, generator=g).to(DEV)
    pos = torch.randn(50, d, generator=g).to(DEV)
    idd = ids.to(DEV)
    x = torch.empty(B * S, d, device=DEV)
    x_c = torch.empty(B * S, d, device=DEV, dtype=torch.bfloat16)
    L.check(lib.vct_embed_fwd(idd.data_ptr(), S + 1, E.data_ptr(), pos.data_ptr(), x.data_ptr(), x_c.data_ptr(), L.BF16,
                              B, S, d, V, 0, 0.0, None, 0, stream()))
    torch.cuda.synchronize()
    want = E[idd[:, :S]] + pos[:S]
    torch.testing.assert_close(x.view(B, S, d), want, rtol=0, atol=0)
    torch.testing.assert_close(x_c.float().view(B, S, d), want, rtol=1e-2, atol=1e-2)
    dx = torch.randn(B * S, d, generator=g).to(DEV)
    dE = torch.zeros(V, d, device=DEV)
    L.check(lib.vct_embed_bwd(idd.data_ptr(), S + 1, dx.data_ptr(), dE.data_ptr(), B, S, d, V, 0, 0.0, None, 0, stream()))
    torch.cuda.synchronize()
    Ew = E.clone().requires_grad_(True)
    torch.nn.functional.embedding(idd[:, :S], Ew, padding_idx=0).backward(dx.view(B, S, d))
    torch.testing.assert_close(dE, Ew.grad, rtol=1e-5, atol=1e-6)
    assert dE[0].abs().sum().item() == 0.0      # padding_idx row
    # sparse form (data-parallel exchange): masked rows, then the same scatter with p = 0 reproduces the dense gradient
    rng = torch.tensor([5, 2], dtype=torch.int64, device=DEV)
    rows = torch.full((B * S, d), float("nan"), device=DEV)
    L.check(lib.vct_embed_bwd_rows(dx.data_ptr(), rows.data_ptr(), B, S, d, 0.3, rng.data_ptr(), 1, stream()))
    dE_direct, dE_sparse = torch.zeros(V, d, device=DEV), torch.zeros(V, d, device=DEV)
    L.check(lib.vct_embed_bwd(idd.data_ptr(), S + 1, dx.data_ptr(), dE_direct.data_ptr(), B, S, d, V, 0, 0.3, rng.data_ptr(), 1, stream()))
    L.check(lib.vct_embed_bwd(idd.data_ptr(), S + 1, rows.data_ptr(), dE_sparse.data_ptr(), B, S, d, V, 0, 0.0, None, 0, stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(dE_sparse, dE_direct, rtol=1e-6, atol=1e-6)
    keep = (rows != 0).float().mean().item()
    assert abs(keep - 0.7) < 0.05
    # deterministic scatter (what the data-parallel ranks run on the gathered rows): equals the atomic scatter up to fp32
    # summation order, is bit-reproducible, and leaves untouched rows alone
    for (Bd, Sd, dd, Vd) in ((B, S, d, V), (96, 20, 768, 30522), (13, 7, 512, 1000)):
        gi = torch.Generator().manual_seed(Bd + Sd)
        ids2 = torch.randint(0, Vd, (Bd, Sd + 1), generator=gi)
        ids2[:, 0] = 101 % Vd                                   # one long segment ([CLS] of every caption)
        ids2[::3, 3:] = 0                                       # pad tails
        r2 = torch.randn(Bd * Sd, dd, generator=gi).to(DEV)
        i2 = ids2.to(DEV)
        want2 = torch.zeros(Vd, dd, device=DEV)
        L.check(lib.vct_embed_bwd(i2.data_ptr(), Sd + 1, r2.data_ptr(), want2.data_ptr(), Bd, Sd, dd, Vd, 0, 0.0, None, 0, stream()))
        outs = []
        for _ in range(2):
            got2 = torch.zeros(Vd, dd, device=DEV)
            got2[1] = 7.0                                       # id 1 may or may not occur: only checked when untouched
            keys = torch.zeros(Bd * Sd, dtype=torch.int32, device=DEV)
            L.check(lib.vct_embed_bwd_det(i2.data_ptr(), Sd + 1, r2.data_ptr(), got2.data_ptr(), Bd, Sd, dd, Vd, 0, keys.data_ptr(), stream()))
            torch.cuda.synchronize()
            outs.append(got2)
        assert torch.equal(outs[0], outs[1])
        used = torch.zeros(Vd, dtype=torch.bool, device=DEV)
        used[i2[:, :Sd].reshape(-1)] = True
        used[0] = False
        torch.testing.assert_close(outs[0][used], want2[used], rtol=1e-5, atol=1e-5)
        if not bool(used[1]):
            assert float((outs[0][1] - 7.0).abs().sum()) == 0.0
        assert float(outs[0][0].abs().sum()) == 0.0             # padding_idx row never written
    # vct_embed_zero clears exactly the touched rows
    dirty = torch.ones(V, d, device=DEV)
    L.check(lib.vct_embed_zero(idd.data_ptr(), S + 1, dirty.data_ptr(), B, S, d, V, stream()))
    torch.cuda.synchronize()
    touched = torch.zeros(V, dtype=torch.bool, device=DEV)
    touched[idd[:, :S].reshape(-1)] = True
    assert float(dirty[touched].abs().sum()) == 0.0 and float((dirty[~touched] - 1).abs().sum()) == 0.0


def test_adam_reads_bf16_gradients(lib):
    """Data-parallel buckets exchanged in bf16: vct_adam with g_dtype = BF16 equals the fp32 kernel fed the rounded values."""
    n = 2048
    g = torch.Generator().manual_seed(4)
    p0, gr = torch.randn(n, generator=g), torch.randn(n, generator=g)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, 0.1, 0.001], device=DEV)
    outs = []
    for dt in (L.BF16, L.F32):
        p = p0.clone().to(DEV)
        m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        g16 = gr.to(DEV, torch.bfloat16)
        gin = g16 if dt == L.BF16 else g16.float()
        L.check(lib.vct_adam(p.data_ptr(), gin.data_ptr(), dt, m.data_ptr(), v.data_ptr(), None, n, hyper.data_ptr(), 0.5, stream()))
        torch.cuda.synchronize()
        outs.append((p.clone(), m.clone(), v.clone()))
    for a, b in zip(*outs):
        torch.testing.assert_close(a, b, rtol=0, atol=0)


@pytest.mark.parametrize("alpha", [0.5, 1.0])
@pytest.mark.parametrize("B,S,V", [(4, 7, 211), (8, 20, 30522), (3, 5, 1000)])
def test_sce_forward_backward_against_oracle(lib, alpha, B, S, V):
    from oracle import vct_oracle as O
    g = torch.Generator().manual_seed(V + B)
    N = B * S
    Vp = (V + 7) // 8 * 8
    logits = (torch.randn(N, V, generator=g) * 2.0)
    ids = torch.randint(1, V, (B, S + 1), generator=g)
    ids[1, 4:] = 0
    zz = logits.clone().requires_grad_(True)
    want = O.sce_loss(zz, ids[:, 1:].reshape(-1), alpha, 1 - alpha, 0)
    want.backward()
    z = torch.zeros(N, Vp)
    z[:, :V] = logits
    z, idd = z.to(DEV), ids.to(DEV)
    loss = torch.zeros(1, device=DEV)
    parts = torch.empty(N, 2, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    dl = torch.full((N, Vp), float("nan"), device=DEV)
    for _ in range(2):
        L.check(lib.vct_sce(z.data_ptr(), Vp, idd.data_ptr(), S + 1, B, S, V, alpha, 1 - alpha, 0, loss.data_ptr(),
                            parts.data_ptr(), counter.data_ptr(), dl.data_ptr(), L.F32, Vp, None, stream()))
    torch.cuda.synchronize()
    assert abs(loss.item() - want.item()) < 2e-5 * max(1.0, abs(want.item()))
    torch.testing.assert_close(dl[:, :V].cpu(), zz.grad, rtol=2e-4, atol=1e-8)
    assert dl[:, V:].abs().sum().item() == 0.0
    # upstream scaling + bf16 gradient output
    up = torch.tensor([0.25], device=DEV)
    dlb = torch.empty((N, Vp), device=DEV, dtype=torch.bfloat16)
    L.check(lib.vct_sce(z.data_ptr(), Vp, idd.data_ptr(), S + 1, B, S, V, alpha, 1 - alpha, 0, None, None, None,
                        dlb.data_ptr(), L.BF16, Vp, up.data_ptr(), stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(dlb[:, :V].float().cpu(), 0.25 * zz.grad, rtol=1e-2, atol=1e-7)


@pytest.mark.parametrize("alpha", [0.5, 1.0])
@pytest.mark.parametrize("B,S,V", [(8, 20, 30522), (3, 5, 1000)])
def test_sce_bf16_logits_against_oracle(lib, alpha, B, S, V):
    """vct_sce_typed with the logits STORED in bf16 (what the training plans feed it): loss and gradient of the oracle
    evaluated on the same bf16-rounded logits (fp32 arithmetic on both sides: tolerances as for fp32 logits)."""
    from oracle import vct_oracle as O
    g = torch.Generator().manual_seed(V + B + 1)
    N, Vp = B * S, (V + 7) // 8 * 8
    logits = (torch.randn(N, V, generator=g) * 2.0).to(torch.bfloat16)
    ids = torch.randint(1, V, (B, S + 1), generator=g)
    ids[1, 3:] = 0
    zz = logits.float().clone().requires_grad_(True)
    want = O.sce_loss(zz, ids[:, 1:].reshape(-1), alpha, 1 - alpha, 0)
    want.backward()
    z = torch.full((N, Vp), 7.0, dtype=torch.bfloat16)          # padding columns hold garbage: must be ignored
    z[:, :V] = logits
    z, idd = z.to(DEV), ids.to(DEV)
    loss = torch.zeros(1, device=DEV)
    parts = torch.empty(N, 2, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    dl = torch.full((N, Vp), float("nan"), device=DEV, dtype=torch.bfloat16)
    for _ in range(2):
        L.check(lib.vct_sce_typed(z.data_ptr(), L.BF16, Vp, idd.data_ptr(), S + 1, B, S, V, alpha, 1 - alpha, 0, loss.data_ptr(),
                                  parts.data_ptr(), counter.data_ptr(), dl.data_ptr(), L.BF16, Vp, None, stream()))
    torch.cuda.synchronize()
    assert abs(loss.item() - want.item()) < 2e-5 * max(1.0, abs(want.item()))
    torch.testing.assert_close(dl[:, :V].float().cpu(), zz.grad, rtol=1e-2, atol=1e-7)      # bf16 rounding of the gradient
    assert dl[:, V:].float().abs().sum().item() == 0.0
    dl32 = torch.full((N, Vp), float("nan"), device=DEV)
    L.check(lib.vct_sce_typed(z.data_ptr(), L.BF16, Vp, idd.data_ptr(), S + 1, B, S, V, alpha, 1 - alpha, 0, None, None, None,
                              dl32.data_ptr(), L.F32, Vp, None, stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(dl32[:, :V].cpu(), zz.grad, rtol=2e-4, atol=1e-8)


def test_sce_staged_kernel_still_covers_unaligned_rows(lib):
    """Row strides that are not multiples of 8 run on the shared-memory kernel (the register kernel needs 16-byte groups)."""
    from oracle import vct_oracle as O
    B, S, V, ld = 3, 4, 1001, 1004
    g = torch.Generator().manual_seed(5)
    N = B * S
    logits = torch.randn(N, V, generator=g)
    ids = torch.randint(1, V, (B, S + 1), generator=g)
    zz = logits.clone().requires_grad_(True)
    want = O.sce_loss(zz, ids[:, 1:].reshape(-1), 0.5, 0.5, 0)
    want.backward()
    z = torch.zeros(N, ld)
    z[:, :V] = logits
    z, idd = z.to(DEV), ids.to(DEV)
    loss = torch.zeros(1, device=DEV)
    parts = torch.empty(N, 2, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    dl = torch.full((N, ld), float("nan"), device=DEV)
    L.check(lib.vct_sce(z.data_ptr(), ld, idd.data_ptr(), S + 1, B, S, V, 0.5, 0.5, 0, loss.data_ptr(), parts.data_ptr(),
                        counter.data_ptr(), dl.data_ptr(), L.F32, ld, None, stream()))
    torch.cuda.synchronize()
    assert abs(loss.item() - want.item()) < 2e-5 * max(1.0, abs(want.item()))
    torch.testing.assert_close(dl[:, :V].cpu(), zz.grad, rtol=2e-4, atol=1e-8)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,ld", [(1280, 768, 768), (100, 30522, 30528), (7, 40, 48)])
def test_colsum(lib, dtype, M, N, ld):
    X = torch.randn(M, ld, generator=torch.Generator().manual_seed(M)).to(DEV, dtype)
    out = torch.empty(N, device=DEV)
    partials = torch.empty(lib.vct_colsum_workspace_floats(M, N), device=DEV)
    counters = torch.zeros(256, dtype=torch.int32, device=DEV)
    for _ in range(2):
        L.check(lib.vct_colsum(X.data_ptr(), L.BF16 if dtype == torch.bfloat16 else L.F32, ld, M, N, out.data_ptr(),
                               partials.data_ptr(), counters.data_ptr(), stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(out, X[:, :N].float().sum(0), rtol=1e-4, atol=1e-3)
    assert int(counters.sum().item()) == 0


def test_adam_matches_torch_optim_and_writes_bf16_shadow(lib):
    from oracle import vct_oracle as O
    n = 4096 + 64
    g = torch.Generator().manual_seed(21)
    p0 = torch.randn(n, generator=g)
    p = p0.clone().to(DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    shadow = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    hyper = torch.tensor([1e-4, 0.9, 0.999, 1e-8, 0.0, 0.0, 0.0, 0.0], device=DEV)
    rng = torch.tensor([1, 0], dtype=torch.int64, device=DEV)
    tp = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([tp], lr=1e-4, betas=(0.9, 0.999))
    for step in range(1, 4):
        gr = torch.randn(n, generator=g)
        tp.grad = gr.clone()
        opt.step()
        L.check(lib.vct_step_tick(rng.data_ptr(), hyper.data_ptr(), stream()))
        L.check(lib.vct_adam(p.data_ptr(), gr.to(DEV).data_ptr(), L.F32, m.data_ptr(), v.data_ptr(), shadow.data_ptr(), n,
                             hyper.data_ptr(), 1.0, stream()))
        torch.cuda.synchronize()
        torch.testing.assert_close(p.cpu(), tp.detach(), rtol=1e-6, atol=1e-7)
    assert int(rng[1].item()) == 3 and hyper[5].item() == 3.0
    torch.testing.assert_close(shadow.float(), p.to(torch.bfloat16).float(), rtol=0, atol=0)


def test_prep_frames_and_cast(lib):
    B, T, Din = 3, 12, 512
    feats = torch.randn(B, T, Din, generator=torch.Generator().manual_seed(4)).to(DEV)
    out = torch.empty(B * (T + 1), Din, device=DEV)
    L.check(lib.vct_prep_frames(feats.data_ptr(), out.data_ptr(), L.F32, B, T, Din, stream()))
    torch.cuda.synchronize()
    o = out.view(B, T + 1, Din)
    torch.testing.assert_close(o[:, 1:], feats, rtol=0, atol=0)
    torch.testing.assert_close(o[:, 0], feats.mean(1), rtol=1e-6, atol=1e-6)
    src = torch.randn(1003, device=DEV)
    dst = torch.empty(1003, device=DEV, dtype=torch.bfloat16)
    L.check(lib.vct_cast(src.data_ptr(), dst.data_ptr(), L.BF16, 1003, stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(dst, src.to(torch.bfloat16), rtol=0, atol=0)


def test_argmax_append_ties_and_end_flags(lib):
    B, V, ld = 4, 30522, 30528
    logits = torch.randn(B, ld, generator=torch.Generator().manual_seed(8)).to(DEV)
    logits[0, 77] = 50.0
    logits[0, 5000] = 50.0                       # tie -> lowest index (torch.max semantics, Q11)
    logits[1, 102] = 60.0                        # end token
    logits[2, V:] = 1e9                          # padding columns must be ignored
    ys = torch.zeros(B, 6, dtype=torch.int64, device=DEV)
    ended = torch.zeros(B, dtype=torch.int32, device=DEV)
    n_ended = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.check(lib.vct_argmax_append(logits.data_ptr(), ld, B, V, ys.data_ptr(), 6, 2, 102, ended.data_ptr(),
                                  n_ended.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert ys[1:, 2].tolist() == logits[1:, :V].argmax(1).tolist()
    assert ys[0, 2].item() == 77 and ys[1, 2].item() == 102
    assert ended.tolist() == [0, 1, 0, 0] and n_ended.item() == 1


# ---- tcgen05 GEMM (TMA + UMMA, bf16 operands, fp32 TMEM accumulators) ---------------------------------
TC_SHAPES = [(128, 256, 64), (128, 64, 128), (256, 512, 768), (1280, 768, 768), (832, 2304, 768), (200, 136, 72),
             (64, 96, 40), (1280, 2048, 768), (768, 2048, 1280), (300, 30522, 96), (1280, 768, 30522),
             (30522, 768, 200), (5000, 2000, 136)]


@pytest.mark.parametrize("a_trans,b_trans", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_gemm_tcgen05_layouts(lib, a_trans, b_trans, M, N, K):
    """All four operand-major combinations (forward, dgrad, wgrad), ragged M/N/K tails handled by TMA
    zero fill, against an fp64 product of the same bf16-rounded operands.  fp32 accumulation: the only
    error is summation order, ~ sqrt(K) * 2^-24 * |terms|."""
    if K % 8:
        K = (K + 7) // 8 * 8
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    lda = (M + 7) // 8 * 8 if a_trans else K
    ldb = (N + 7) // 8 * 8 if b_trans else K
    Ad = torch.zeros((K, lda) if a_trans else (M, lda), dtype=torch.bfloat16, device=DEV)
    Bd = torch.zeros((K, ldb) if b_trans else (N, ldb), dtype=torch.bfloat16, device=DEV)
    if a_trans: Ad[:, :M] = A.t().to(DEV)
    else: Ad[:, :K] = A.to(DEV)
    if b_trans: Bd[:, :N] = B.t().to(DEV)
    else: Bd[:, :K] = B.to(DEV)
    Ar = (Ad[:, :M].t() if a_trans else Ad[:, :K]).double()
    Br = (Bd[:, :N].t() if b_trans else Bd[:, :K]).double()
    ldc = (N + 3) // 4 * 4
    Cc, _ = run_gemm(lib, Ad, Bd, a_trans, b_trans, M, N, K, impl=L.GEMM_TCGEN05, ldc=ldc)
    want = Ar @ Br.t()
    got = Cc[:, :N].double()
    assert torch.isfinite(got).all()
    torch.testing.assert_close(got, want, rtol=1e-4, atol=5e-4 * math.sqrt(K))


def test_gemm_tcgen05_epilogues_match_simt(lib):
    M, N, K = 300, 2048, 768
    g = torch.Generator().manual_seed(77)
    A = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    B = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    addend = torch.randn(M, N, generator=g).to(DEV)
    table = torch.randn(13, N, generator=g).to(DEV)
    rng = torch.tensor([9, 4], dtype=torch.int64, device=DEV)
    for kw in (dict(bias=bias, addend=addend, row_table=table, row_period=13, c2_dtype=L.BF16),
               dict(bias=bias, act=L.ACT_GELU_FWD, c2_dtype=L.BF16, drop_p=0.3, rng_state=rng, site=5, c_dtype=L.BF16),
               dict(act=L.ACT_GELU_BWD, aux=torch.randn(M, N, generator=g).to(DEV, torch.bfloat16), drop_p=0.3,
                    rng_state=rng, site=5, c_dtype=L.BF16)):
        cd = kw.pop("c_dtype", L.F32)
        c_s, c2_s = run_gemm(lib, A, B, 0, 0, M, N, K, c_dtype=cd, impl=L.GEMM_SIMT, **kw)
        c_t, c2_t = run_gemm(lib, A, B, 0, 0, M, N, K, c_dtype=cd, impl=L.GEMM_TCGEN05, **kw)
        tol = dict(rtol=2e-2, atol=2e-2) if cd == L.BF16 else dict(rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(c_t.float(), c_s.float(), **tol)
        if c2_s is not None:
            torch.testing.assert_close(c2_t.float(), c2_s.float(), rtol=2e-2, atol=2e-2)


# ---- fused in-projection + attention kernel (attn_fused.cu) vs projection GEMM + stand-alone core -------------
def _run_self_flavour(lib, x, w, b, key_pad, B, Lq, d, H, causal, impl, drop_p=0.0, rng=None):
    from vct import lib as VL
    m = VL.MhaArgs()
    qkv = torch.full((B * Lq, 3 * d), float("nan"), device=DEV, dtype=torch.bfloat16)
    o = torch.full((B * Lq, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    probs = torch.zeros(B, H, Lq, Lq, device=DEV)
    m.B, m.L, m.Lk, m.d, m.H, m.dtype = B, Lq, Lq, d, H, L_BF16
    m.x, m.w_in, m.b_in = x.data_ptr(), w.data_ptr(), b.data_ptr()
    m.qkv, m.o, m.key_pad = qkv.data_ptr(), o.data_ptr(), key_pad.data_ptr()
    m.drop_p, m.rng_state, m.site = drop_p, (rng.data_ptr() if rng is not None else None), 321
    m.probs = probs.data_ptr()
    m.gemm_impl = impl
    fn = lib.vct_attn_dec_self_fwd if causal else lib.vct_attn_enc_self_fwd
    L_check(fn(C.byref(m), stream()))
    torch.cuda.synchronize()
    return qkv, o, probs


L_BF16 = L.BF16
L_check = L.check


@pytest.mark.parametrize("B,Lq,d,H,causal", [(64, 13, 768, 8, 0), (64, 20, 768, 8, 1), (7, 20, 768, 8, 1), (16, 33, 768, 8, 0),
                                             (64, 13, 512, 8, 0), (5, 20, 512, 8, 1), (3, 64, 512, 8, 1)])
def test_fused_self_attention_matches_composition(lib, B, Lq, d, H, causal):
    # impl = TCGEN05 runs the fused tensor-core kernel (attn_fused.cu; sequences up to 32), impl = SIMT the composition
    g = torch.Generator().manual_seed(B + Lq + d)
    x = torch.randn(B * Lq, d, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(3 * d, d, generator=g) * 0.04).to(DEV, torch.bfloat16)
    b = (torch.randn(3 * d, generator=g) * 0.1).to(DEV)
    key_pad = torch.zeros(B, Lq, dtype=torch.uint8)
    for bb in range(1, B):
        key_pad[bb, Lq - (bb % 5):] = 1 if bb % 5 else 0
    key_pad = key_pad.to(DEV)
    for p in (0.0, 0.3):
        rng = torch.tensor([77, 5], dtype=torch.int64, device=DEV)
        qf, of, pf = _run_self_flavour(lib, x, w, b, key_pad, B, Lq, d, H, causal, L.GEMM_TCGEN05, p, rng)
        qs, os_, ps = _run_self_flavour(lib, x, w, b, key_pad, B, Lq, d, H, causal, L.GEMM_SIMT, p, rng)
        assert torch.isfinite(of.float()).all() and torch.isfinite(qf.float()).all()
        torch.testing.assert_close(qf.float(), qs.float(), rtol=2e-2, atol=2e-2)      # both bf16-rounded projections
        # both attend over bf16-rounded q/k/v; the fused kernel also rounds the probabilities to bf16 for the P V MMA
        torch.testing.assert_close(pf, ps, rtol=5e-2, atol=5e-3)
        torch.testing.assert_close(of.float(), os_.float(), rtol=5e-2, atol=3e-2)


def _run_cross_flavour(lib, x, mem, w, b, B, Lq, Lk, d, H, impl, drop_p=0.0, rng=None):
    from vct import lib as VL
    # K/V of the memory pre-projected (kv_ready = 1), exactly like the training plan does for every decoder layer
    kv = (mem.float() @ w[d:].float().t() + b[d:]).to(torch.bfloat16).contiguous()
    m = VL.MhaArgs()
    q = torch.full((B * Lq, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    o = torch.full((B * Lq, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    probs = torch.zeros(B, H, Lq, Lk, device=DEV)
    m.B, m.L, m.Lk, m.d, m.H, m.dtype = B, Lq, Lk, d, H, L_BF16
    m.x, m.mem, m.w_in, m.b_in = x.data_ptr(), mem.data_ptr(), w.data_ptr(), b.data_ptr()
    m.qkv, m.kv, m.kv_ready, m.o = q.data_ptr(), kv.data_ptr(), 1, o.data_ptr()
    m.drop_p, m.rng_state, m.site = drop_p, (rng.data_ptr() if rng is not None else None), 654
    m.probs = probs.data_ptr()
    m.gemm_impl = impl
    L_check(lib.vct_attn_dec_cross_fwd(C.byref(m), stream()))
    torch.cuda.synchronize()
    return q, o, probs


@pytest.mark.parametrize("B,Lq,Lk,d,H", [(64, 20, 13, 768, 8), (7, 20, 13, 768, 8), (16, 20, 33, 768, 8), (128, 20, 33, 768, 8),
                                         (5, 20, 13, 512, 8), (9, 25, 20, 768, 8), (64, 12, 13, 768, 8)])
def test_fused_cross_attention_matches_composition(lib, B, Lq, Lk, d, H):
    """decoder cross flavour: q projection + attention in one tensor-core kernel (TCGEN05) against projection GEMM +
    stand-alone core (SIMT); Lq <= 16 (last case) is outside the fused kernel's coverage and checks the fallback."""
    g = torch.Generator().manual_seed(B + Lq + Lk + d)
    x = torch.randn(B * Lq, d, generator=g).to(DEV, torch.bfloat16)
    mem = torch.randn(B * Lk, d, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(3 * d, d, generator=g) * 0.04).to(DEV, torch.bfloat16)
    b = (torch.randn(3 * d, generator=g) * 0.1).to(DEV)
    for p in (0.0, 0.3):
        rng = torch.tensor([78, 9], dtype=torch.int64, device=DEV)
        qf, of, pf = _run_cross_flavour(lib, x, mem, w, b, B, Lq, Lk, d, H, L.GEMM_TCGEN05, p, rng)
        qs, os_, ps = _run_cross_flavour(lib, x, mem, w, b, B, Lq, Lk, d, H, L.GEMM_SIMT, p, rng)
        assert torch.isfinite(of.float()).all() and torch.isfinite(qf.float()).all()
        torch.testing.assert_close(qf.float(), qs.float(), rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(pf, ps, rtol=5e-2, atol=5e-3)
        torch.testing.assert_close(of.float(), os_.float(), rtol=5e-2, atol=3e-2)


@pytest.mark.parametrize("B,H,Lq,Lk,dh,causal", [(64, 8, 20, 20, 96, 1), (64, 8, 13, 13, 96, 0), (64, 8, 20, 13, 96, 0), (7, 8, 20, 20, 96, 1),
                                                 (16, 8, 20, 33, 96, 0), (5, 8, 20, 13, 64, 0), (9, 8, 25, 25, 64, 1)])
def test_attention_backward_tcgen05_matches_simt(lib, B, H, Lq, Lk, dh, causal):
    """bf16 attention backward: the tensor-core kernel (attn_fused.cu) against the SIMT kernel (forced by passing explicit
    gradient batch strides, which the tensor-core path does not take), with key padding, dropout and the fused
    in-projection bias gradient."""
    g = torch.Generator().manual_seed(B + 7 * Lq + Lk)
    d = H * dh
    q = torch.randn(B, Lq, H, dh, generator=g).to(DEV, torch.bfloat16)
    k = torch.randn(B, Lk, H, dh, generator=g).to(DEV, torch.bfloat16)
    v = torch.randn(B, Lk, H, dh, generator=g).to(DEV, torch.bfloat16)
    do = torch.randn(B, Lq, H, dh, generator=g).to(DEV, torch.bfloat16)
    key_pad = None
    if Lq == Lk:
        key_pad = torch.zeros(B, Lk, dtype=torch.uint8)
        for b in range(1, B):
            key_pad[b, Lk - (b % 6):] = 1 if b % 6 else 0
        key_pad = key_pad.to(DEV)
    rng = torch.tensor([91, 3], dtype=torch.int64, device=DEV)
    outs = []
    for force_simt in (False, True):
        dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
        dbias = torch.full((3 * d,), float("nan"), device=DEV)
        part = torch.empty(B, 3 * d, device=DEV)
        cnt = torch.zeros(H, dtype=torch.int32, device=DEV)
        a = L.AttnArgs()
        a.B, a.H, a.Lq, a.Lk, a.dh, a.dtype = B, H, Lq, Lk, dh, L.BF16
        a.q, a.q_ld, a.k, a.k_ld, a.v, a.v_ld = q.data_ptr(), d, k.data_ptr(), d, v.data_ptr(), d
        a.key_pad = key_pad.data_ptr() if key_pad is not None else None
        a.causal, a.scale = causal, 1.0 / math.sqrt(dh)
        a.drop_p, a.rng_state, a.site = 0.3, rng.data_ptr(), 17
        a.d_o, a.do_ld, a.dq, a.dq_ld, a.dk, a.dk_ld, a.dv, a.dv_ld = do.data_ptr(), d, dq.data_ptr(), d, dk.data_ptr(), d, dv.data_ptr(), d
        a.dbias, a.dbias_partials, a.dbias_counters = dbias.data_ptr(), part.data_ptr(), cnt.data_ptr()
        if force_simt:
            a.dq_bs, a.dk_bs, a.dv_bs = Lq * d, Lk * d, Lk * d
        L.check(lib.vct_attn_bwd(C.byref(a), stream()), "bwd")
        torch.cuda.synchronize()
        assert int(cnt.abs().sum()) == 0                       # counters reset themselves
        outs.append((dq.float(), dk.float(), dv.float(), dbias.clone()))
    (dq_t, dk_t, dv_t, db_t), (dq_s, dk_s, dv_s, db_s) = outs
    for t_, s_ in ((dq_t, dq_s), (dk_t, dk_s), (dv_t, dv_s)):
        assert torch.isfinite(t_).all()
        torch.testing.assert_close(t_, s_, rtol=3e-2, atol=3e-2)
        assert float((t_ - s_).norm() / s_.norm()) < 1e-2
    assert float((db_t - db_s).norm() / db_s.norm()) < 5e-3
    # the bias gradient is the column sum of the (unrounded) gradients
    ref = torch.cat([dq_s.reshape(-1, d).sum(0), dk_s.reshape(-1, d).sum(0), dv_s.reshape(-1, d).sum(0)])
    assert float((db_s - ref).norm() / ref.norm()) < 2e-2


@pytest.mark.parametrize("impl", [L.GEMM_SIMT, L.GEMM_TCGEN05])
def test_gemm_saved_gelu_factor_pair_matches_recompute_pair(lib, impl):
    """ACT_GELU_FWD_F stores f = gelu'(z) * dropmask in the forward epilogue and ACT_MUL_AUX multiplies by it in backward
    (what the training plans use); the result must equal the ACT_GELU_FWD / ACT_GELU_BWD pair that recomputes gelu'
    and the mask from z."""
    M, N, K = 1280, 2048, 256
    g = torch.Generator().manual_seed(5)
    A = (torch.randn(M, K, generator=g) * 0.2).to(DEV, torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    G = (torch.randn(M, K, generator=g) * 0.2).to(DEV, torch.bfloat16)          # upstream operand of the backward GEMM
    W = torch.randn(N, K, generator=g).to(DEV, torch.bfloat16)
    rng = torch.tensor([11, 4], dtype=torch.int64, device=DEV)
    kw = dict(bias=bias, c2_dtype=L.BF16, drop_p=0.3, rng_state=rng, site=9, c_dtype=L.BF16, impl=impl)
    z, h1 = run_gemm(lib, A, B, 0, 0, M, N, K, act=L.ACT_GELU_FWD, **kw)
    f, h2 = run_gemm(lib, A, B, 0, 0, M, N, K, act=L.ACT_GELU_FWD_F, **kw)
    torch.testing.assert_close(h2.float(), h1.float(), rtol=1e-2, atol=1e-2)
    zz = z.float().clone().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    keep = (h1.float() != 0) | (z.float().abs() < 1e-3)                          # dropped <=> h == 0 (up to z ~ 0)
    torch.testing.assert_close(f.float() * keep, zz.grad / 0.7 * keep * (f.float() != 0), rtol=3e-2, atol=3e-2)
    d1, _ = run_gemm(lib, G, W, 0, 0, M, N, K, act=L.ACT_GELU_BWD, aux=z, drop_p=0.3, rng_state=rng, site=9, c_dtype=L.BF16, impl=impl)
    d2, _ = run_gemm(lib, G, W, 0, 0, M, N, K, act=L.ACT_MUL_AUX, aux=f, c_dtype=L.BF16, impl=impl)
    assert torch.isfinite(d2.float()).all()
    assert float((d2.float() - d1.float()).norm() / d1.float().norm()) < 1e-2


@pytest.mark.parametrize("B,T,Din,S1", [(64, 12, 512, 21), (3, 5, 24, 8), (2, 70, 512, 81)])
@pytest.mark.parametrize("with_masks", [False, True])
def test_stage_inputs_matches_the_torch_copies(lib, B, T, Din, S1, with_masks):
    """vct_stage_inputs == feats copy, [B,T+1] frame mask with an unpadded global column, ids copy, ids[:, :-1] == pad
    (or the caller's mask); either half can be skipped."""
    g = torch.Generator().manual_seed(B + T)
    feats = torch.randn(B, T, Din, generator=g).to(DEV)
    vid = (torch.rand(B, T, generator=g) < 0.3).to(DEV) if with_masks else None
    ids = torch.randint(0, 5, (B, S1), generator=g).to(DEV)
    tok = (torch.rand(B, S1 - 1, generator=g) < 0.5).to(DEV) if with_masks else None
    f2 = torch.full((B, T, Din), float("nan"), device=DEV)
    v2 = torch.full((B, T + 1), 7, dtype=torch.uint8, device=DEV)
    i2 = torch.full((B, S1), -1, dtype=torch.int64, device=DEV)
    t2 = torch.full((B, S1 - 1), 7, dtype=torch.uint8, device=DEV)
    L.check(lib.vct_stage_inputs(feats.data_ptr(), f2.data_ptr(), vid.data_ptr() if vid is not None else None, v2.data_ptr(),
                                 B, T, Din, ids.data_ptr(), i2.data_ptr(), tok.data_ptr() if tok is not None else None,
                                 t2.data_ptr(), S1, 0, stream()))
    torch.cuda.synchronize()
    assert torch.equal(f2, feats) and torch.equal(i2, ids)
    want_v = torch.zeros(B, T + 1, dtype=torch.uint8, device=DEV)
    if vid is not None:
        want_v[:, 1:] = vid.to(torch.uint8)
    assert torch.equal(v2, want_v)
    assert torch.equal(t2, (tok if tok is not None else ids[:, :-1] == 0).to(torch.uint8))
    # token half only: the feature buffers stay untouched
    f2.fill_(1.0), v2.fill_(9)
    L.check(lib.vct_stage_inputs(None, f2.data_ptr(), None, v2.data_ptr(), B, T, Din, ids.data_ptr(), i2.data_ptr(), None,
                                 t2.data_ptr(), S1, 3, stream()))
    torch.cuda.synchronize()
    assert float(f2.min()) == 1.0 and int(v2.min()) == 9
    assert torch.equal(t2, (ids[:, :-1] == 3).to(torch.uint8))
