"""Epoch-end caption generation sharded over the data-parallel ranks.

The reference evaluates on rank 0 only, one video per ``greedy_decode`` call (json ``data.eval.batch_size`` = 1), while
the other ranks wait at a barrier (train.py:171-185, 244-256): at 8 GPUs 7 of them idle for the whole pass.  Greedy
decoding has no cross-sample dependence, so here every rank decodes a disjoint slice of the videos in large batches with
the K/V-cached decoder and the per-video strings are exchanged once (``all_gather_object``: a few hundred KB).  The result
is the ``vid2result`` dict the reference builds at train.py:177-180; scoring it (COCO / Java) stays with the caller.

No collective on the data path: replicas only."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Videos of rank ``rank``: a strided slice, so that every rank gets the same count +- 1 and (for length-sorted
    datasets) the same mix of lengths.  The shards are disjoint and cover range(n) exactly (no padding duplicates, unlike
    DistributedSampler, because every video must be scored exactly once)."""
    return list(range(rank, n, world))


def _strip(caption: str) -> str:
    return caption.replace("[CLS]", "").replace("[SEP]", "")       # eval.py:141


@torch.no_grad()
def sharded_greedy_eval(model, dataset, max_len: int = 30, batch_size: int = 256, rank: Optional[int] = None,
                        world: Optional[int] = None, group=None, with_masks: bool = True) -> Dict[str, str]:
    """``{video id: predicted caption}`` for every video of ``dataset`` (a ``vct.data.PackedCaptionDataset``), identical on
    every rank.  ``model`` is the MMT4Caption core (``model.module`` under DDP); ``with_masks=False`` reproduces
    predict_video.py (no masks, no eval fast path), True reproduces eval.py:140 / train.py eval_epoch (masks passed)."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    if rank is None:
        rank = dist.get_rank(group) if distributed else 0
    if world is None:
        world = dist.get_world_size(group) if distributed else 1
    was_training = getattr(model, "training", False)
    model.eval()
    mode = dataset.mode
    dataset.mode = "by_video"
    mine = shard_indices(len(dataset.video_feat_list), rank, world)
    # Batches hold videos of ONE frame count: the reference decodes one video per call (json data.eval.batch_size = 1), and
    # a padded batch would not reproduce it -- cross-attention is never masked, so padded frames of a ragged
    # batch are attended to.  (CLIP4Clip uni_12 features are all 12 frames: one group.)
    lens = dataset.lengths[0].tolist()
    groups: Dict[int, List[int]] = {}
    for v in mine:
        groups.setdefault(int(lens[v]), []).append(v)
    local: Dict[str, str] = {}
    try:
        for _, members in sorted(groups.items()):
            for i in range(0, len(members), batch_size):
                feats, masks, _, vids = dataset.batch(members[i:i + batch_size])
                caps = model.greedy_decode(feats, masks if with_masks else None, max_len=max_len)
                local.update(zip(vids, (_strip(c) for c in caps)))
    finally:
        dataset.mode = mode
        if was_training:
            model.train()
    if world == 1 or not distributed:
        return local
    parts: List[Optional[Dict[str, str]]] = [None] * world
    dist.all_gather_object(parts, local, group=group)
    merged: Dict[str, str] = {}
    for p in parts:
        merged.update(p)
    return merged


def make_coco_inputs(vid2result: Dict[str, str], video2caption: Dict[str, Sequence[str]]):
    """(gts, samples, ids) in the layout the reference's scorer takes (eval.py:20-39 make_coco_sample, used at
    train.py:182): samples[vid] = [{"image_id", "caption"}] for every prediction, gts[vid] = one such dict per ground-truth
    caption of EVERY annotated video."""
    samples = {vid: [{"image_id": vid, "caption": pred}] for vid, pred in vid2result.items()}
    ids = list(vid2result.keys())
    gts = {vid: [{"image_id": vid, "caption": c} for c in caps] for vid, caps in video2caption.items()}
    return gts, samples, ids
