"""Host-side rows N2 / N3 / N4 on CPU: the packed, device-resident data path against the reference's
own loader on a synthetic MSR-VTT-layout dataset, batched tokenisation against caption-by-caption encoding, and the
rank-sharded evaluation under a world_size-2 gloo group."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _dataset(tmp_path, ragged=False):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import json
    import numpy as np
    rng = np.random.default_rng(1)
    out = str(tmp_path)
    os.makedirs(os.path.join(out, "feats"), exist_ok=True)
    videos, sentences = [], []
    for i in range(11):
        T = int(rng.integers(3, 13)) if ragged else 12
        np.save(os.path.join(out, "feats", f"video{i}.npy"), rng.standard_normal((T, 512)).astype(np.float32))
        videos.append({"video_id": f"video{i}", "split": "train"})
        for _ in range(3):
            sentences.append({"video_id": f"video{i}", "caption": " ".join(f"w{int(t)}" for t in rng.integers(1000, 30522, int(rng.integers(2, 9))))})
    ann = os.path.join(out, "ann.json")
    with open(ann, "w") as f:
        json.dump({"videos": videos, "sentences": sentences}, f)
    return {"feat_dir": [os.path.join(out, "feats")], "annotation_path": ann, "split_mode": "train", "mode": "by_caption",
            "batch_size": 4, "_debug": False, "_debug_num": 400, "dataset": "msrvtt"}


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("mode", ["by_caption", "by_video"])
def test_packed_loader_matches_reference_loader(tmp_path, mode, ragged):
    from vct.data import PackedCaptionDataset, PackedLoader
    cfg = dict(_dataset(tmp_path, ragged), mode=mode)
    ds = PackedCaptionDataset(cfg["feat_dir"], cfg["annotation_path"], split_type="train", mode=mode)
    ours = list(PackedLoader(ds, 4))
    if os.path.isdir(REF):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_vct_ref_dataloader", os.path.join(REF, "dataloader.py"))
        rdl = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(rdl)
        rds = rdl.MSRVTT_Dataset(cfg["feat_dir"], cfg["annotation_path"], split_type="train", mode=mode)
        ref = list(torch.utils.data.DataLoader(rds, batch_size=4, collate_fn=rdl.collate_fn, shuffle=False))
        assert len(ref) == len(ours) and len(rds) == len(ds)
        assert rds.video2caption == ds.video2caption
    else:                                            # GPU box: the same contract restated (dataloader.py:237-274, 500-504)
        ref = []
        for i in range(0, len(ds), 4):
            items = [ds[j] for j in range(i, min(len(ds), i + 4))]
            tmax = max(it[0][0].shape[0] for it in items)
            x = torch.zeros(len(items), tmax, 512)
            m = torch.ones(len(items), tmax, dtype=torch.bool)
            for k, it in enumerate(items):
                x[k, :it[0][0].shape[0]] = it[0][0]
                m[k, :it[0][0].shape[0]] = False
            ref.append(([x], [m], tuple(it[1] for it in items), tuple(it[2] for it in items)))
    for (rf, rm, rc, rv), (of, om, oc, ov) in zip(ref, ours):
        assert len(rf) == len(of) == 1
        assert torch.equal(rf[0], of[0]) and torch.equal(rm[0], om[0]) and rm[0].dtype == om[0].dtype == torch.bool
        assert tuple(rc) == tuple(oc) and tuple(rv) == tuple(ov)


def test_pretokenized_batches_equal_cap_preprocessor(tmp_path, tokenizer_dir):
    from vct.data import build_packed_dataloader
    from model.CapPreprocessor import CapPreprocessor
    cp = CapPreprocessor(tokenizer_dir, device=torch.device("cpu"))
    cfg = _dataset(tmp_path)
    ds, loader, sampler = build_packed_dataloader(cfg, multi_gpu=False, cap_preprocessor=cp)
    assert sampler is None and loader.pretokenized
    loader.shuffle = False
    n = 0
    for (f, m, ids, vids), i in zip(loader, range(0, len(ds), 4)):
        caps = [ds.cap_vid_list[j][0] for j in range(i, min(len(ds), i + 4))]
        want = torch.full((len(caps), max(len(cp.tokenizer.encode(c)) for c in caps)), cp.pad_id, dtype=torch.long)
        for k, c in enumerate(caps):                                     # the reference's loop, model/CapPreprocessor.py:24-33
            e = cp.tokenizer.encode(c)
            want[k, :len(e)] = torch.tensor(e)
        assert torch.equal(ids, want)
        ids2, mask2 = cp(caps)
        assert torch.equal(ids2, want) and torch.equal(mask2, want == cp.pad_id)
        n += 1
    assert n == len(loader)


def test_distributed_sampler_shards_like_the_reference(tmp_path):
    """build_packed_dataloader(multi_gpu=True) uses torch's DistributedSampler like dataloader.py:523-525: two ranks see
    disjoint halves of one seeded permutation, re-drawn by set_epoch."""
    from vct.data import build_packed_dataloader
    cfg = _dataset(tmp_path)
    parts = []
    for r in range(2):
        ds, loader, sampler = build_packed_dataloader(cfg, multi_gpu=True, rank=r, world_size=2)
        sampler.set_epoch(3)
        ws = torch.utils.data.DistributedSampler(ds, num_replicas=2, rank=r, shuffle=True)
        ws.set_epoch(3)
        got = [v for b in loader for v in b[3]]
        assert got == [ds.cap_vid_list[i][1][0].stem for i in ws]
        parts.append(list(ws))
    assert len(set(parts[0]) | set(parts[1])) == len(ds)


class _FakeModel:
    """greedy_decode stand-in: the caption is a pure function of the video's features, so any sharding / batching of the
    videos must give the same vid -> caption map."""
    training = False

    def eval(self):
        return self

    def greedy_decode(self, feats, masks, max_len=30):
        x = feats[0]
        keep = (~masks[0]).float().unsqueeze(-1) if masks is not None else torch.ones_like(x[..., :1])
        return [f"[CLS]c{int(round(float(v) * 1000))}[SEP]" for v in (x * keep).sum((1, 2))]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _eval_worker(rank, world, port, cfg, ret):
    sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200"))
    from vct.data import PackedCaptionDataset
    from vct.evaluate import sharded_greedy_eval
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = PackedCaptionDataset(cfg["feat_dir"], cfg["annotation_path"], split_type="train", mode="by_video")
    res = sharded_greedy_eval(_FakeModel(), ds, max_len=30, batch_size=3)
    ret[rank] = res
    dist.destroy_process_group()


def test_sharded_eval_world2_gloo_equals_single_process(tmp_path):
    from vct.data import PackedCaptionDataset
    from vct.evaluate import sharded_greedy_eval, shard_indices, make_coco_inputs
    cfg = _dataset(tmp_path, ragged=True)
    ds = PackedCaptionDataset(cfg["feat_dir"], cfg["annotation_path"], split_type="train", mode="by_caption")
    single = sharded_greedy_eval(_FakeModel(), ds, batch_size=1, rank=0, world=1)       # the reference: one video per call
    assert ds.mode == "by_caption" and len(single) == 11 and all(not v.startswith("[CLS]") for v in single.values())
    assert sorted(shard_indices(11, 0, 2) + shard_indices(11, 1, 2)) == list(range(11))
    ret = mp.Manager().dict()
    mp.spawn(_eval_worker, args=(2, _free_port(), cfg, ret), nprocs=2, join=True)
    assert dict(ret[0]) == dict(ret[1]) == single
    gts, samples, ids = make_coco_inputs(single, ds.video2caption)
    assert set(ids) == set(single) and all(len(gts[v]) == 3 for v in ids) and samples[ids[0]][0]["caption"] == single[ids[0]]
