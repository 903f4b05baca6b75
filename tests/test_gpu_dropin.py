"""The call surface of the reference's train.py / eval.py / predict_video.py exercised on the GPU with the drop-in
package (the reference tree itself is not available on the GPU box): DistributedDataParallel wrap, raw caption
strings through the tokenizer, torch.optim.Adam + CosineAnnealingLR, loss all-reduce, no_grad validation through
``model.module``, greedy decoding to strings, checkpoint save / ``load_state_dict(strict=False)``, and the
``layer.forward`` rebinding of predict_video.py:126-130 (cross-attention maps in ``layer.mha``)."""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_reference_call_surface(tokenizer_dir, tmp_path):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from model.MMT4Caption import MMT4Caption
    from vct.synthetic import shipped_model_config
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        torch.manual_seed(666)
        cfg = shipped_model_config(tokenizer_dir, dropout=0.0)
        model = MMT4Caption(cfg, device=dev).to(dev)                       # train.py:210
        model.vct_precision = "fp32"
        model.mode("caption")                                              # train.py:211
        ddp = DDP(model, device_ids=[0], output_device=0)                  # train.py:218
        opt = torch.optim.Adam(filter(lambda p: p.requires_grad, ddp.parameters()), lr=1e-4, betas=(0.9, 0.999))
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=8, eta_min=1e-5)
        g = torch.Generator().manual_seed(3)
        B = 6
        feats = [torch.randn(B, 12, 512, generator=g)]
        masks = [torch.zeros(B, 12, dtype=torch.bool)]
        captions = tuple(" ".join(f"w{int(t)}" for t in torch.randint(1000, 30522, (int(n),), generator=g))
                         for n in torch.randint(3, 10, (B,), generator=g))
        ddp.train()
        ddp.module.mode("caption")
        losses = []
        for _ in range(4):                                                  # train.py:119-131
            v_feats = [i.to(dev) for i in feats]
            v_masks = [i.to(dev) for i in masks]
            loss = ddp(v_feats, v_masks, captions)
            opt.zero_grad()
            loss.backward()
            opt.step()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
            losses.append(loss.item())
        sched.step()
        assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses
        ddp.eval()                                                          # train.py:151-168
        with torch.no_grad():
            vloss = ddp.module([i.to(dev) for i in feats], [i.to(dev) for i in masks], captions).item()
        assert vloss < losses[0]
        caps = ddp.module.greedy_decode([feats[0][:2].to(dev)], None, max_len=8)     # eval.py:140 / train.py:199-205
        assert isinstance(caps, list) and len(caps) == 2 and all(isinstance(c, str) for c in caps)
        path = str(tmp_path / "ckpt.pth")
        torch.save(ddp.module.state_dict(), path)                           # utils.py:53-60
        fresh = MMT4Caption(cfg, device=dev).to(dev)
        fresh.vct_precision = "fp32"
        fresh.mode("caption")
        state = fresh.load_state_dict(torch.load(path, map_location=dev), strict=False)   # eval.py:149-151
        assert not state.missing_keys and not state.unexpected_keys
        fresh.eval()
        with torch.no_grad():
            vloss2 = fresh([i.to(dev) for i in feats], [i.to(dev) for i in masks], captions).item()
        assert abs(vloss2 - vloss) < 1e-5 * abs(vloss)

        # predict_video.py:43-79,126-130: every decoder layer's forward is rebound to a function that records `mha`
        def attn_forward(self, tgt, memory, **kw):
            raise AssertionError("the patched forward must not be needed: the fused path provides layer.mha itself")
        for layer in fresh.cap_decoder.decoder.layers:
            layer.forward = types.MethodType(attn_forward, layer)
        out = fresh.greedy_decode([feats[0][:1].to(dev)], max_len=6)
        assert len(out) == 1
        for layer in fresh.cap_decoder.decoder.layers:
            assert layer.mha.shape[0] == 1 and layer.mha.shape[2] == 13 and layer.mha.shape[1] >= 1
            torch.testing.assert_close(layer.mha.sum(-1), torch.ones_like(layer.mha.sum(-1)), rtol=1e-4, atol=1e-4)
    finally:
        dist.destroy_process_group()
