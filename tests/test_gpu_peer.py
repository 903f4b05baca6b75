"""NVLink peer-memory collectives (csrc/peer_comm.cu).  World 1 runs on any GPU box; the 2-rank cases need two GPUs in
one process (gpurun --gpus 2) and are skipped otherwise.  The N-process path (cudaIpc handles exchanged over
torch.distributed) is exercised by tools/ddp_check.py and bench.py under torchrun."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from vct import lib as L  # noqa: E402


def _make(world, nbytes, ctas=8):
    lib = L.load()
    hs, bases = [], []
    for r in range(world):
        with torch.cuda.device(r):
            h = C.c_void_p()
            L.check(lib.vct_comm_create(r, world, nbytes, ctas, C.byref(h)), "vct_comm_create")
            hs.append(h)
            bases.append(int(lib.vct_comm_base(h)))
    if world > 1:
        pb = (C.c_void_p * world)(*bases)
        pd = (C.c_int * world)(*range(world))
        for r in range(world):
            L.check(lib.vct_comm_connect_in_process(hs[r], pb, pd), "vct_comm_connect_in_process")
    return lib, hs, bases


def _view(base, n, dtype, dev):
    from vct.peer import _Raw
    if dtype == torch.bfloat16:
        return torch.as_tensor(_Raw(base, (n,), "<i2", None), device=f"cuda:{dev}").view(torch.bfloat16)
    return torch.as_tensor(_Raw(base, (n,), "<f4", None), device=f"cuda:{dev}")


def test_world1_allreduce_is_identity_and_graph_capturable():
    lib, hs, bases = _make(1, 1 << 20)
    x = _view(bases[0], 4096, torch.bfloat16, 0)
    ref = torch.randn(4096, device="cuda:0").to(torch.bfloat16)
    x.copy_(ref)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        L.check(lib.vct_peer_allreduce_bf16(hs[0], 0, 4096, 65536, 0, st.cuda_stream), "allreduce")
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            L.check(lib.vct_peer_allreduce_bf16(hs[0], 0, 4096, 65536, 0, st.cuda_stream), "allreduce")
            L.check(lib.vct_peer_allgather(hs[0], 8192, 1024, 1, st.cuda_stream), "allgather")
        for _ in range(3):
            g.replay()
    torch.cuda.synchronize()
    assert torch.equal(x, ref) and lib.vct_comm_status(hs[0]) == 0
    assert lib.vct_peer_allreduce_bf16(hs[0], 8, 4096, 65536, 0, None) < 0   # misaligned offset is rejected
    assert lib.vct_peer_allreduce_bf16(hs[0], 0, 4096, 4096, 0, None) < 0    # staging overlapping the data is rejected
    lib.vct_comm_destroy(hs[0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
@pytest.mark.parametrize("n", [8, 4096, 1000 * 8, 3_000_000])
def test_two_rank_allreduce_matches_fp32_sum_and_is_identical_on_both(n):
    lib, hs, bases = _make(2, 16 << 20)
    g = torch.Generator().manual_seed(n)
    src = [torch.randn(n, generator=g).to(torch.bfloat16) for _ in range(2)]
    off = 512
    views = [_view(bases[r] + off, n, torch.bfloat16, r) for r in range(2)]
    streams = [torch.cuda.Stream(device=r) for r in range(2)]
    for rep in range(3):                                   # repeated calls: the epoch counters keep the barriers apart
        for r in range(2):
            views[r].copy_(src[r].to(f"cuda:{r}"))
        for r in range(2):
            torch.cuda.synchronize(r)
        for r in range(2):
            with torch.cuda.device(r):
                L.check(lib.vct_peer_allreduce_bf16(hs[r], off, n, 8 << 20, 0, streams[r].cuda_stream), "allreduce")
        for r in range(2):
            torch.cuda.synchronize(r)
        want = (src[0].float() + src[1].float()).to(torch.bfloat16)
        got0, got1 = views[0].cpu(), views[1].cpu()
        assert torch.equal(got0, got1)
        assert torch.equal(got0, want)
    assert lib.vct_comm_status(hs[0]) == 0 and lib.vct_comm_status(hs[1]) == 0
    for h in hs:
        lib.vct_comm_destroy(h)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_rank_allgather():
    lib, hs, bases = _make(2, 4 << 20)
    slot = 12345 * 16
    n = slot // 4
    parts = [torch.randn(n, generator=torch.Generator().manual_seed(r)) for r in range(2)]
    views = [_view(bases[r], 2 * n, torch.float32, r) for r in range(2)]
    for r in range(2):
        views[r].zero_()
        views[r][r * n:(r + 1) * n].copy_(parts[r].to(f"cuda:{r}"))
        torch.cuda.synchronize(r)
    for r in range(2):
        with torch.cuda.device(r):
            L.check(lib.vct_peer_allgather(hs[r], 0, slot, 2, torch.cuda.current_stream(r).cuda_stream), "allgather")
    want = torch.cat(parts)
    for r in range(2):
        torch.cuda.synchronize(r)
        assert torch.equal(views[r].cpu(), want)
    for h in hs:
        lib.vct_comm_destroy(h)
