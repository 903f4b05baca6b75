"""The autograd drop-in path (EncoderFn / DecoderFn -- what the reference's train.py and DistributedDataParallel drive)
under the conditions round 1 did not test: dropout p > 0, gradient accumulation, a caller-supplied key-padding
mask, outputs that outlive the next forward, and a workspace reused between forward and backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_tiny  # noqa: E402

DEV = torch.device("cuda")


def build(name="tiny_a", precision="fp32", gemm="simt", dropout=0.0):
    from model.MMEncoder import MultiModalEncoder
    from model.CapDecoder import CapDecoder
    cfg, sd, ins, outs, grads = load_tiny(name)
    enc = MultiModalEncoder([cfg["Din"]], cfg["d"], cfg["nhead"], cfg["F"], cfg["L_enc"], dropout, "gelu", "avg", True,
                            "encoding", False, DEV)
    dec = CapDecoder(cfg["L_dec"], cfg["d"], cfg["nhead"], cfg["F"], dropout, cfg["V"], 0, cfg["alpha"], None, "gelu", DEV)
    enc.load_state_dict({k[len("video_encoder."):]: v for k, v in sd.items() if k.startswith("video_encoder.")})
    dec.load_state_dict({k[len("cap_decoder."):]: v for k, v in sd.items() if k.startswith("cap_decoder.")})
    enc.to(DEV), dec.to(DEV)
    for m in (enc, dec):
        m.vct_precision, m.vct_gemm = precision, gemm
    return cfg, enc, dec, ins, outs, grads


def full_model(tokenizer_dir, dropout, precision="fp32", gemm="simt"):
    from model.MMT4Caption import MMT4Caption
    from vct.synthetic import shipped_model_config
    torch.manual_seed(666)
    cfg = shipped_model_config(tokenizer_dir, embed_dim=64, enc_layers=1, dec_layers=2, nhead=2, feedforward=96,
                               dropout=dropout, modal_shape=(32,))
    m = MMT4Caption(cfg, device=DEV).to(DEV)
    m.vct_precision, m.vct_gemm = precision, gemm
    m.mode("caption")
    return m


def small_batch(B=6, T=5, Din=32, S1=9, seed=5):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, Din, generator=g)
    tok = torch.randint(1000, 30522, (B, S1), generator=g)
    tok[:, 0], tok[:, -1] = 101, 102
    return x.to(DEV), torch.zeros(B, T, dtype=torch.bool, device=DEV), tok.to(DEV)


def test_dropout_masks_change_between_training_forwards_and_match_in_backward(tokenizer_dir):
    """the autograd path never advanced the dropout step, so every training step applied the same
    masks.  Now: (1) consecutive ``model(...)`` calls on identical inputs give different losses (fresh masks); (2) the
    debug mask entry shows different keep patterns for the two steps; (3) backward regenerates the forward's masks: the
    analytic gradient equals a central finite difference of the loss evaluated AT THE SAME dropout step."""
    from vct import lib as L
    model = full_model(tokenizer_dir, dropout=0.3)
    model.train()
    eng = model._engine()
    x, vm, tok = small_batch()
    l1 = model([x], [vm], tok)
    step1 = eng._host_step
    l2 = model([x], [vm], tok)
    step2 = eng._host_step
    assert step2 == step1 + 1
    assert abs(float(l1) - float(l2)) > 1e-4, (float(l1), float(l2))
    masks = []
    for st in (step1, step2):
        eng.set_rng_step(st)
        m = torch.zeros(4096, dtype=torch.uint8, device=DEV)
        L.check(eng.lib.vct_dropout_mask(m.data_ptr(), m.numel(), 0.3, eng.rng_state.data_ptr(), 1,
                                         torch.cuda.current_stream().cuda_stream))
        masks.append(m.clone())
    assert abs(float(masks[0].float().mean()) - 0.7) < 0.03
    assert int((masks[0] != masks[1]).sum()) > 1000
    # (3) gradient vs finite difference at a pinned step, with an unrelated forward between forward and backward
    model.zero_grad(set_to_none=True)
    loss = model([x], [vm], tok)
    pinned = eng._host_step
    x2, vm2, tok2 = small_batch(B=4, S1=7, seed=9)            # other shape: other workspace, but it advances the RNG step
    _ = model([x2], [vm2], tok2)
    loss.backward()
    p = model.cap_decoder.decoder.layers[1].linear2.weight
    g = p.grad.clone()
    gen = torch.Generator().manual_seed(1)
    direction = torch.randn(p.shape, generator=gen).to(DEV)
    direction /= direction.norm()
    eps = 2e-2
    vals = []
    with torch.no_grad():
        for sgn in (+1.0, -1.0):
            p.add_(sgn * eps * direction)
            eng._host_step = pinned - 1                       # the next training forward draws `pinned` again
            eng._dev_step = None
            vals.append(float(model([x], [vm], tok)))
            assert eng._host_step == pinned
            p.sub_(sgn * eps * direction)
    fd = (vals[0] - vals[1]) / (2 * eps)
    an = float((g * direction).sum())
    assert abs(fd - an) <= 2e-2 * max(abs(fd), abs(an)) + 2e-5, (fd, an)


def test_ranks_draw_different_dropout_seeds():
    from vct.engine import CaptionEngine
    s0 = CaptionEngine._rank_seed(666)
    assert s0 == 666                                          # no process group: rank 0
    import torch.distributed as dist
    assert not dist.is_initialized()


def test_gradient_accumulation_over_two_backwards_matches_oracle():
    """with p.grad already set (no set_to_none zeroing) a second backward must ADD its gradient.
    p.grad of the first backward is a view of the gradient arena that the second backward overwrites -- it is moved
    out of the arena first.  Checked against golden gradients: g(batch) + g(batch) = 2 g."""
    cfg, enc, dec, ins, outs, grads = build("tiny_a")
    enc.train(), dec.train()
    x, vm, ids = ins["feats"].to(DEV), ins["vid_pad"].to(DEV), ins["ids"].to(DEV)
    for _ in range(2):
        memory, _, _ = enc([x], [vm])
        _, loss = dec(memory, ids, ids == 0)
        loss.backward()
    for pre, mod in (("video_encoder.", enc), ("cap_decoder.", dec)):
        for k, p in mod.named_parameters():
            want = 2.0 * grads[pre + k].to(DEV)
            err = float((p.grad - want).norm() / (want.norm() + 1e-12))
            assert err <= 2e-3, (pre + k, err)
    # and zero_grad(set_to_none=False) followed by one backward gives g again
    for mod in (enc, dec):
        for p in mod.parameters():
            p.grad.zero_()
    memory, _, _ = enc([x], [vm])
    _, loss = dec(memory, ids, ids == 0)
    loss.backward()
    k = "cap_decoder.generator.weight"
    err = float((dec.generator.weight.grad - grads[k].to(DEV)).norm() / grads[k].norm())
    assert err <= 2e-3, err


def test_outputs_are_fresh_tensors_and_stale_workspace_raises():
    """SURVEY 8b "outputs freshly allocated": memory / logits returned by one forward keep their values after another
    forward of the same shape; backward through the FIRST forward then raises (its saved activations are gone)
    instead of returning wrong gradients."""
    cfg, enc, dec, ins, outs, _ = build("tiny_a")
    enc.train(), dec.train()
    x, vm, ids = ins["feats"].to(DEV), ins["vid_pad"].to(DEV), ins["ids"].to(DEV)
    mem1, _, _ = enc([x], [vm])
    logits1, loss1 = dec(mem1, ids, ids == 0)
    keep_mem, keep_logits = mem1.detach().clone(), logits1.detach().clone()
    mem2, _, _ = enc([x * 0.5], [vm])
    logits2, loss2 = dec(mem2, ids, ids == 0)
    assert torch.equal(mem1.detach(), keep_mem) and torch.equal(logits1, keep_logits)
    assert not torch.equal(mem2.detach(), keep_mem)
    with pytest.raises(RuntimeError, match="reused by another forward"):
        loss1.backward()
    loss2.backward()                                           # the latest forward is still differentiable
    assert dec.generator.weight.grad is not None


def test_decoder_uses_the_padding_mask_it_is_given():
    """model/CapDecoder.py:43-52: the key-padding mask is the ARGUMENT's [:, :-1], not `tgt == pad_id`.  Passing an
    all-False mask for a padded batch must equal the oracle run with that mask; passing None equals `tgt == pad`."""
    from oracle import vct_oracle as O
    cfg, enc, dec, ins, outs, _ = build("tiny_a")
    _, sd, _, _, _ = load_tiny("tiny_a")
    enc.eval(), dec.eval()
    x, vm, ids = ins["feats"].to(DEV), ins["vid_pad"].to(DEV), ins["ids"].to(DEV)
    mem = torch.from_numpy(outs["memory"]).to(DEV)            # the golden memory (the eval fast path would alter padded rows)
    with torch.no_grad():
        lg_default, _ = dec(mem, ids, None)
        lg_same, _ = dec(mem, ids, ids == 0)
        lg_nomask, _ = dec(mem, ids, torch.zeros_like(ids, dtype=torch.bool))
    assert torch.equal(lg_default, lg_same)
    torch.testing.assert_close(lg_same.cpu(), torch.from_numpy(outs["logits"]), rtol=0, atol=2e-4)
    no_pad = torch.zeros(ins["ids"].shape[0], ins["ids"].shape[1] - 1, dtype=torch.bool)
    want = O.generator(sd, O.decoder_hidden(sd, torch.from_numpy(outs["memory"]), ins["ids"][:, :-1], no_pad, cfg["nhead"]))
    torch.testing.assert_close(lg_nomask.cpu(), want, rtol=0, atol=2e-4)
    assert float((lg_nomask - lg_same).abs().max()) > 1e-3     # the padded batch really depends on the mask


def test_workspace_cache_is_bounded(monkeypatch):
    """batches built by caption length have a new S almost every step; the per-shape workspaces (and
    the plans / graphs stored in them) are kept in a small LRU instead of growing without bound."""
    monkeypatch.setenv("VCT_WS_CACHE", "3")
    cfg, enc, dec, ins, outs, _ = build("tiny_a")
    from vct.engine import CaptionEngine
    from model._engine import module_dims
    eng = CaptionEngine(enc, dec, dims=module_dims(enc, dec), device=torch.device("cuda", 0), precision="fp32")
    seen = [eng.workspace(4, 5, S, True) for S in range(3, 10)]
    assert len([k for k in eng._ws if k[0] != "decode"]) <= 3
    assert eng.workspace(4, 5, 9, True) is seen[-1]            # most recent ones are still cached
    assert eng.workspace(4, 5, 3, True) is not seen[0]         # the oldest was evicted and is rebuilt on demand
