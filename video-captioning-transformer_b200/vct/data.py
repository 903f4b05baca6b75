"""Device-resident data path for the caption task.

The reference's loader (dataloader.py:354-532) runs ``num_workers=0`` and does one ``np.load`` per sample per step, pads
the batch on the host, and tokenises caption by caption inside the model.  Once the GPU step takes ~1.7 ms that loader IS
the step.  Here a split is packed ONCE:

  * every feature file of the split -> one ``[N, Tmax, Din]`` fp32 tensor + ``lengths [N]`` on the device
    (MSR-VTT train: 6.5 k videos x 12 x 512 x 4 B = 160 MB; CLIP4Clip ``uni_12`` features are all 12 frames),
  * every caption -> one ``[C, Lmax]`` int64 id table (one batched tokenizer call), resident on the device,

and a batch is two gathers by index (``index_select``), produced on the device with no host work per sample.  The objects
mirror the reference's: ``PackedCaptionDataset`` has ``video_feat_list``, ``cap_vid_list``, ``video2caption``, ``mode``,
``__len__`` / ``__getitem__`` with the reference's item layout, ``build_packed_dataloader(data_cfg, multi_gpu)`` returns
``(data_iter, dataloader, sampler)`` like ``dataloader.build_dataloader`` (dataloader.py:507-532), and the loader yields the
reference's collate tuple ``(feat_ts, feat_mask_ts, batch_captions, batch_vids)`` (dataloader.py:500-504) -- with
``batch_captions`` either the raw strings or (``pretokenize=True``) the id tensor ``MMT4Caption.forward`` also accepts.
Sharding across ranks uses ``torch.utils.data.DistributedSampler`` itself, so every rank sees exactly the samples the
reference's loader would give it (dataloader.py:523-525)."""
from __future__ import annotations

import json
import pathlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import DistributedSampler


def _load_feature(path) -> torch.Tensor:
    """[T, E] fp32; a file stored [E, T] is transposed (dataloader.py:381-385)."""
    t = torch.from_numpy(np.load(str(path))).to(torch.float32)
    return t.transpose(0, 1) if t.shape[0] > t.shape[1] else t


class PackedCaptionDataset:
    """One split of MSR-VTT / MSVD with every feature file loaded once into a padded tensor."""

    def __init__(self, video_feat_dirs: Sequence[str], annotation_file: str, dataset: str = "msrvtt", split_type: str = "train",
                 mode: str = "by_caption", debug: bool = False, debug_num: int = 400, device=None):
        if split_type.lower() in ("val", "validate"):
            split_type = "validate"
        if mode not in ("by_caption", "by_video"):
            raise ValueError(mode)
        self.split_type, self.mode, self.dataset = split_type, mode, dataset
        self.annotation_file, self.video_feat_dirs = annotation_file, list(video_feat_dirs)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        # one tuple of per-modality paths per video, in the reference's (glob) order (dataloader.py:368-372)
        globs = [list(pathlib.Path(d).glob("*.npy")) for d in self.video_feat_dirs]
        self.video_feat_list: List[Tuple[pathlib.Path, ...]] = list(zip(*globs))
        self.vids: List[str] = [v[0].stem for v in self.video_feat_list]
        self._vid_index: Dict[str, int] = {v: i for i, v in enumerate(self.vids)}
        self.cap_vid_list, self.video2caption = self._make_cap_vid_list()
        if debug:
            self.cap_vid_list = self.cap_vid_list[:debug_num]
        self._cap_video = torch.tensor([self._vid_index[p[0].stem] for _, p in self.cap_vid_list], dtype=torch.long)
        self._pack()
        self.ids_table: Optional[torch.Tensor] = None          # [C, Lmax] int64 after pretokenize()
        self.ids_len: Optional[torch.Tensor] = None

    # ---- annotations (dataloader.py:410-436 MSR-VTT json, :470-491 MSVD text) ---------------------------------------
    def _make_cap_vid_list(self):
        video2caption: Dict[str, List[str]] = {}
        if self.dataset == "msrvtt":
            with open(self.annotation_file, encoding="utf-8") as f:
                ann = json.load(f)
            video2split = {v["video_id"]: v["split"] for v in ann["videos"]}
            for cap in ann["sentences"]:
                if video2split[cap["video_id"]] != self.split_type:
                    continue
                video2caption.setdefault(cap["video_id"], []).append(cap["caption"])
        else:
            with open(self.annotation_file) as f:
                for line in f.readlines():
                    vid = line.split(" ")[0]
                    cap = " ".join(line.split(" ")[1:]).replace("\n", "")
                    video2caption.setdefault(vid, []).append(cap)
        video2path = {p[0].stem: p for p in self.video_feat_list}
        cap_vid_list = [(cap, video2path[video]) for video, caps in video2caption.items() for cap in caps]
        return cap_vid_list, video2caption

    # ---- packing ------------------------------------------------------------------------------------------------
    def _pack(self) -> None:
        n_modal = len(self.video_feat_dirs)
        self.feats: List[torch.Tensor] = []          # per modality [N, Tmax, E]
        self.lengths: List[torch.Tensor] = []        # per modality [N]
        for m in range(n_modal):
            items = [_load_feature(paths[m]) for paths in self.video_feat_list]
            tmax = max((t.shape[0] for t in items), default=1)
            E = items[0].shape[1] if items else 1
            packed = torch.zeros((len(items), tmax, E), dtype=torch.float32)
            lens = torch.zeros(len(items), dtype=torch.long)
            for i, t in enumerate(items):
                packed[i, :t.shape[0]] = t
                lens[i] = t.shape[0]
            self.feats.append(packed.to(self.device))
            self.lengths.append(lens.to(self.device))

    def pretokenize(self, cap_preprocessor) -> None:
        """Tokenise every caption of the split once (``CapPreprocessor.encode_host``: one batched tokenizer call)."""
        ids = cap_preprocessor.encode_host([c for c, _ in self.cap_vid_list])
        self.ids_len = (ids != cap_preprocessor.pad_id).sum(1).to(self.device)
        self.ids_table = ids.to(self.device)
        self.pad_id = cap_preprocessor.pad_id

    # ---- reference item API (dataloader.py:377-397) -----------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.cap_vid_list) if self.mode == "by_caption" else len(self.video_feat_list)

    def __getitem__(self, index):
        if self.mode == "by_caption":
            caption, v = self.cap_vid_list[index][0], int(self._cap_video[index])
        else:
            caption, v = "", index
        feats = [f[v, :int(l[v])] for f, l in zip(self.feats, self.lengths)]
        return feats, caption, self.vids[v]

    # ---- batched access ---------------------------------------------------------------------------------------------
    def batch(self, indices: Sequence[int], pretokenized: bool = False):
        """The reference's collate tuple for the samples ``indices`` (dataloader.py:500-504): feature tensors padded to the
        batch's longest video, masks True = padded, captions (strings, or ids [B, Lbatch] when pretokenized), video ids."""
        idx = torch.as_tensor(indices, dtype=torch.long)
        vsel = (self._cap_video[idx] if self.mode == "by_caption" else idx)
        vdev = vsel.to(self.device)
        feat_ts, mask_ts = [], []
        for f, l in zip(self.feats, self.lengths):
            lens = l.index_select(0, vdev)
            tmax = int(lens.max()) if lens.numel() else 0
            x = f.index_select(0, vdev)[:, :tmax]
            feat_ts.append(x)
            mask_ts.append(torch.arange(tmax, device=self.device)[None, :] >= lens[:, None])
        vids = tuple(self.vids[int(v)] for v in vsel)
        if self.mode != "by_caption":
            caps = tuple("" for _ in vids)
        elif pretokenized:
            if self.ids_table is None:
                raise RuntimeError("call pretokenize(model.cap_preprocessor) first")
            cdev = idx.to(self.device)
            lmax = int(self.ids_len.index_select(0, cdev).max())
            caps = self.ids_table.index_select(0, cdev)[:, :lmax]
        else:
            caps = tuple(self.cap_vid_list[int(i)][0] for i in idx)
        return feat_ts, mask_ts, caps, vids


class PackedLoader:
    """Iterates a ``PackedCaptionDataset`` in batches; same length / drop_last=False semantics as the reference's
    ``DataLoader(data_iter, batch_size, collate_fn, sampler | shuffle)`` (dataloader.py:527-531)."""

    def __init__(self, dataset: PackedCaptionDataset, batch_size: int, sampler=None, shuffle: bool = False,
                 pretokenized: bool = False, seed: int = 0):
        self.dataset, self.batch_size, self.sampler, self.shuffle = dataset, int(batch_size), sampler, shuffle
        self.pretokenized = pretokenized
        self._gen = torch.Generator().manual_seed(seed)

    def _order(self) -> List[int]:
        if self.sampler is not None:
            return list(iter(self.sampler))
        if self.shuffle:
            return torch.randperm(len(self.dataset), generator=self._gen).tolist()
        return list(range(len(self.dataset)))

    def __len__(self) -> int:
        n = len(self.sampler) if self.sampler is not None else len(self.dataset)
        return (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = self._order()
        for i in range(0, len(order), self.batch_size):
            yield self.dataset.batch(order[i:i + self.batch_size], self.pretokenized)


def build_packed_dataloader(data_cfg: dict, multi_gpu: bool, device=None, cap_preprocessor=None, rank: Optional[int] = None,
                            world_size: Optional[int] = None):
    """Drop-in for ``dataloader.build_dataloader(data_cfg, multi_gpu)`` (dataloader.py:507-532): same config keys, same
    return triple.  ``cap_preprocessor`` given -> the captions are tokenised once and the loader yields id tensors."""
    ds = PackedCaptionDataset(data_cfg["feat_dir"], data_cfg["annotation_path"], dataset=data_cfg.get("dataset", "msrvtt"),
                              split_type=data_cfg["split_mode"], mode=data_cfg["mode"], debug=data_cfg.get("_debug", False),
                              debug_num=data_cfg.get("_debug_num", 400), device=device)
    train = data_cfg["split_mode"] == "train"
    sampler = None
    if train and multi_gpu:
        kw = {} if rank is None else {"rank": rank, "num_replicas": world_size}
        sampler = DistributedSampler(ds, shuffle=True, **kw)
    pre = cap_preprocessor is not None and ds.mode == "by_caption"
    if pre:
        ds.pretokenize(cap_preprocessor)
    loader = PackedLoader(ds, data_cfg["batch_size"], sampler=sampler, shuffle=(train and not multi_gpu), pretokenized=pre)
    return ds, loader, sampler
