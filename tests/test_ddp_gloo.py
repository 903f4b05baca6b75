"""world_size-2 gloo test (CPU) of the data-parallel gradient exchange used by CaptionTrainer:
bucketed SUM all-reduce over the flat gradient arena + the 1/world scale folded into Adam gives every
rank the mean gradient and keeps the replicas bit-identical."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "video-captioning-transformer_b200"))
    sys.path.insert(0, root)
    from vct.arena import ParamArena
    from vct.trainer import gradient_buckets, all_reduce_flat
    from oracle import vct_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                         # identical replicas
    enc, dec = torch.nn.Linear(64, 96), torch.nn.Linear(96, 33)
    named = [("video_encoder." + n, p) for n, p in enc.named_parameters()] + \
            [("cap_decoder." + n, p) for n, p in dec.named_parameters()]
    arena = ParamArena(named, torch.device("cpu"))
    buckets = gradient_buckets(arena, ["video_encoder.", "cap_decoder."], max_bytes=4 * 1000)
    # rank-local shard of a global batch (rows [r*B/N, (r+1)*B/N))
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(8, 64, generator=g), torch.randn(8, 33, generator=g)
    xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
    loss = ((dec(torch.relu(enc(xs))) - ys) ** 2).mean()
    grads = torch.autograd.grad(loss, [p for _, p in named])
    for (n, _), gr in zip(named, grads):
        arena.grad_view(n).copy_(gr)
    all_reduce_flat(arena.grad, buckets)
    m, v = torch.zeros_like(arena.p32), torch.zeros_like(arena.p32)
    p_new, _, _ = O.adam_step(arena.p32, arena.grad * (1.0 / world), m, v, 1, 1e-3)
    # single-process reference on the full batch: equal shard sizes => mean of shard grads == global grad
    torch.manual_seed(0)
    enc2, dec2 = torch.nn.Linear(64, 96), torch.nn.Linear(96, 33)
    full = ((dec2(torch.relu(enc2(X))) - Y) ** 2).mean()
    gfull = torch.autograd.grad(full, list(enc2.parameters()) + list(dec2.parameters()))
    ok = True
    for (n, _), gr in zip(named, gfull):
        ok &= torch.allclose(arena.grad_view(n) / world, gr, rtol=1e-5, atol=1e-6)
    gathered = [torch.zeros_like(p_new) for _ in range(world)]
    dist.all_gather(gathered, p_new)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        ret["ok"], ret["same"] = bool(ok), bool(same)
    dist.destroy_process_group()


def test_bucketed_allreduce_gives_mean_gradient_and_identical_replicas():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["ok"], "mean of per-rank gradients != global-batch gradient"
    assert ret["same"], "replicas diverged after the update"


def _sparse_worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "video-captioning-transformer_b200"))
    sys.path.insert(0, root)
    from vct.trainer import gather_embedding_rows, scatter_embedding_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, d, B, S, pad = 50, 8, 3, 5, 0
    g = torch.Generator().manual_seed(10 + rank)                 # every rank holds a DIFFERENT shard
    ids = torch.randint(1, V, (B, S + 1), generator=g)
    ids[rank % B, 2:] = pad                                      # padded positions contribute nothing (padding_idx)
    table = torch.randn(V, d, generator=torch.Generator().manual_seed(3), requires_grad=True)
    w = torch.randn(B, S, d, generator=g)
    loss = (torch.nn.functional.embedding(ids[:, :-1], table, padding_idx=pad) * w).sum()
    dense, = torch.autograd.grad(loss, table)                    # what DDP would all-reduce: the dense [V, d] gradient
    rows = w.reshape(B * S, d).clone()                           # d loss / d embedded row = the sparse form
    all_rows, all_ids = gather_embedding_rows(rows, ids)
    sparse_sum = scatter_embedding_rows(all_rows, all_ids, V, pad)
    dist.all_reduce(dense, op=dist.ReduceOp.SUM)
    # bf16 bucket: cast -> SUM all-reduce -> 1/world in the optimizer; compared with the fp32 exchange
    flat = torch.randn(1000, generator=g)
    f32 = flat.clone()
    dist.all_reduce(f32, op=dist.ReduceOp.SUM)
    b16 = flat.to(torch.bfloat16)
    dist.all_reduce(b16, op=dist.ReduceOp.SUM)
    if rank == 0:
        ret["sparse_equals_dense"] = bool(torch.allclose(sparse_sum, dense, rtol=1e-6, atol=1e-6))
        ret["shapes"] = (tuple(all_rows.shape), tuple(all_ids.shape))
        ret["bf16_rel"] = float((b16.float() - f32).norm() / f32.norm())
    dist.destroy_process_group()


def test_sparse_embedding_exchange_equals_dense_allreduce_and_bf16_bucket_is_close():
    """The N > 1 trainer exchanges the embedding-table gradient as rows + ids (all-gather) and the dense buckets in
    bf16: the first must equal DDP's dense SUM all-reduce exactly (up to fp32 summation order), the second stays within
    bf16 rounding of the fp32 exchange."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sparse_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["sparse_equals_dense"]
    assert ret["shapes"] == ((2 * 15, 8), (2 * 3, 6))
    assert ret["bf16_rel"] < 8e-3


def test_embedding_exchange_mode_needs_the_uniform_shape_promise(monkeypatch):
    """The sparse exchange of the embedding-table gradient has fixed per-rank slots for [B*S, d] rows and [B, S+1] ids, so it
    is only chosen when the caller promises equal (B, S) on every rank (CaptionTrainer(uniform_shapes=True)); ragged batches
    -- what the reference's loader produces, dataloader.py:507-532 -- take the dense, shape-independent exchange.  The
    decision is a pure function of the engine's attributes: checked here without a GPU."""
    import sys
    from types import SimpleNamespace
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video-captioning-transformer_b200"))
    from vct.engine import CaptionEngine
    for k in ("VCT_SPLIT_EMB_ADAM", "VCT_SPARSE_EMB"):
        monkeypatch.delenv(k, raising=False)
    ws = SimpleNamespace(B=64, S=20)
    eng = SimpleNamespace(dims=SimpleNamespace(V=30522), peer=None, PEER_MAX_TOKENS=CaptionEngine.PEER_MAX_TOKENS)
    mode = lambda world: CaptionEngine.emb_mode(eng, ws, world)
    assert mode(1) == "local"
    assert mode(2) == "dense"                                   # no promise: ragged batches must work
    eng.uniform_shapes = True
    assert mode(2) == "sparse" and mode(8) == "sparse"
    assert CaptionEngine.emb_mode(eng, SimpleNamespace(B=512, S=20), 8) == "dense"      # 81920 tokens > one-CTA sort
    eng.dims.V = 50000
    assert mode(2) == "dense"                                   # vocabulary beyond the sort's key width
    eng.dims.V = 30522
    eng.peer = object()                                         # peer all-gather moves 16-byte vectors: even id slots only
    assert CaptionEngine.emb_mode(eng, SimpleNamespace(B=3, S=20), 2) == "dense"
    assert CaptionEngine.emb_mode(eng, SimpleNamespace(B=4, S=20), 2) == "sparse"
    monkeypatch.setenv("VCT_SPARSE_EMB", "0")
    assert mode(2) == "dense"
    monkeypatch.setenv("VCT_SPLIT_EMB_ADAM", "0")
    assert mode(1) == "local-unsplit" and mode(2) == "dense"
