"""GPU end of rows N2-N4: the packed device-resident loader feeding the native trainer, and the batched / sharded
evaluation against one-video-at-a-time greedy decoding (the reference's eval loop, train.py:171-185)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_data_eval import _dataset  # noqa: E402

DEV = torch.device("cuda")


def _model(tokenizer_dir, precision):
    from model.MMT4Caption import MMT4Caption
    from vct.synthetic import shipped_model_config
    torch.manual_seed(666)
    m = MMT4Caption(shipped_model_config(tokenizer_dir, dropout=0.0), device=DEV).to(DEV)
    m.vct_precision = precision
    m.mode("caption")
    return m


def test_packed_loader_feeds_trainer_and_matches_string_captions(tmp_path, tokenizer_dir):
    """One epoch over the packed loader (device-resident features, pre-tokenised ids): every batch's loss equals the loss
    of the same batch given as raw caption strings through the reference-style call model(feats, masks, captions)."""
    from vct.data import build_packed_dataloader
    model = _model(tokenizer_dir, "fp32")
    model.eval()
    cfg = _dataset(tmp_path, ragged=True)
    ds, loader, _ = build_packed_dataloader(cfg, multi_gpu=False, device=DEV, cap_preprocessor=model.cap_preprocessor)
    loader.shuffle = False
    assert ds.feats[0].is_cuda and ds.ids_table.is_cuda
    n = 0
    with torch.no_grad():
        for i, (feats, masks, ids, vids) in zip(range(0, len(ds), 4), loader):
            assert feats[0].is_cuda and ids.is_cuda and ids.dtype == torch.int64
            caps = [ds.cap_vid_list[j][0] for j in range(i, min(len(ds), i + 4))]
            l_ids = float(model(feats, masks, ids))
            l_str = float(model(feats, masks, caps))
            assert abs(l_ids - l_str) <= 1e-6 * abs(l_str), (l_ids, l_str)
            n += 1
    assert n == len(loader)
    from vct.trainer import CaptionTrainer
    model.train()
    tr = CaptionTrainer(model, lr=1e-4, use_graph=False)
    losses = []
    for feats, masks, ids, vids in loader:
        losses.append(float(tr.step(feats[0], masks[0], ids)))
    assert all(torch.isfinite(torch.tensor(losses))) and len(losses) == len(loader)


def test_batched_sharded_eval_equals_per_video_decode(tmp_path, tokenizer_dir):
    from vct.data import PackedCaptionDataset
    from vct.evaluate import sharded_greedy_eval
    model = _model(tokenizer_dir, "fp32")
    cfg = _dataset(tmp_path, ragged=True)
    ds = PackedCaptionDataset(cfg["feat_dir"], cfg["annotation_path"], split_type="train", mode="by_video", device=DEV)
    batched = sharded_greedy_eval(model, ds, max_len=8, batch_size=4)
    model.eval()
    with torch.no_grad():
        for v in range(len(ds)):
            feats, _, vid = ds[v]
            # the reference's loop: batch of one video, masks from the collate (all False for a single video)
            one = model.greedy_decode([feats[0].unsqueeze(0)], [torch.zeros(1, feats[0].shape[0], dtype=torch.bool, device=DEV)],
                                      max_len=8)[0]
            assert batched[vid] == one.replace("[CLS]", "").replace("[SEP]", ""), (vid, batched[vid], one)
    assert len(batched) == len(ds.video_feat_list)


def test_prefetched_steps_equal_direct_steps(tokenizer_dir):
    """trainer.prefetch() only moves the H2D copy off the critical path: losses of prefetched steps are bit-identical to
    steps fed directly, including when the prefetched batch differs from step to step."""
    from vct.synthetic import synth_batch
    from vct.trainer import CaptionTrainer
    batches = [synth_batch(8, 12, 512, 21, seed=100 + k, padded=True) for k in range(4)]
    pinned = [tuple(t.pin_memory() for t in b) for b in batches]
    out = []
    for mode in ("direct", "prefetch"):
        model = _model(tokenizer_dir, "bf16")
        model.train()
        tr = CaptionTrainer(model, lr=1e-4)
        losses = []
        if mode == "prefetch":
            tr.prefetch(*pinned[0])
        for k in range(4):
            x, vm, ids = pinned[k]
            loss = tr.step(x, vm, ids)
            if mode == "prefetch" and k + 1 < 4:
                tr.prefetch(*pinned[k + 1])
            losses.append(float(loss.item()))
        out.append(losses)
    assert out[0] == out[1], out


def test_long_sequences_run_and_over_long_ones_fail_up_front(tokenizer_dir):
    """the fused attention kernels cover sequences up to 64 rows.  Longer ones (the reference truncates neither
    frames nor training captions) now run on the tiled kernels -- checked here at the shipped dims against the oracle --
    and anything beyond their 1024-row limit is rejected before the first launch (not in the middle of an epoch) with a
    message that names the remedy."""
    from oracle import vct_oracle as O
    model = _model(tokenizer_dir, "bf16")
    model.train()
    sd = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 70, 512, generator=g)                      # 70 frames -> memory length 71
    vm = torch.zeros(2, 70, dtype=torch.bool)
    vm[1, 50:] = True
    ids = torch.randint(1000, 30522, (2, 80), generator=g)        # 79 decoder positions
    ids[:, 0], ids[0, -1], ids[1, 40], ids[1, 41:] = 101, 102, 102, 0
    loss = model([x.to(DEV)], [vm.to(DEV)], ids.to(DEV))
    loss.backward()
    _, _, oloss = O.caption_forward(sd, x, vm, ids, 8, 8, 0.5)
    assert abs(float(loss) - float(oloss)) <= 5e-3 * float(oloss), (float(loss), float(oloss))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters() if p.requires_grad)
    x = torch.randn(2, 1100, 512, device=DEV)                     # beyond the tiled kernels (and the reference's 512-row table)
    vm = torch.zeros(2, 1100, dtype=torch.bool, device=DEV)
    with pytest.raises(ValueError, match="sequence too long"):
        model([x], [vm], ids.to(DEV))
    x = torch.randn(2, 12, 512, device=DEV)
    vm = torch.zeros(2, 12, dtype=torch.bool, device=DEV)
    long_ids = torch.randint(1000, 30522, (2, 1100), device=DEV)
    with pytest.raises(ValueError, match="sequence too long"):
        model([x], [vm], long_ids)
    model.eval()
    with pytest.raises(ValueError, match="sequence too long"):
        model.greedy_decode([x], [vm], max_len=1100)
    assert len(model.greedy_decode([x], [vm], max_len=6)) == 2    # the engine is still usable afterwards
