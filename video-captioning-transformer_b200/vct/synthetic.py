"""Synthetic inputs and an offline tokenizer directory for benchmarks / smoke runs.

There is no network on the build or GPU boxes, so ``bert-base-uncased`` (the tokenizer the
shipped configs name, configs/*.json ``model.tokenizer``) cannot be fetched.  This module
writes a BertTokenizer directory with the same special-token ids and vocabulary size
(30522) so that ``MMT4Caption(cfg)`` constructs offline, and generates the synthetic
``[B, T, 512]`` frame features + random token ids that BASELINE.json's configs name.
"""
from __future__ import annotations

import json
import os

import torch

VOCAB_SIZE = 30522
PAD_ID, UNK_ID, CLS_ID, SEP_ID, MASK_ID = 0, 100, 101, 102, 103


def make_tokenizer_dir(path: str, vocab_size: int = VOCAB_SIZE) -> str:
    os.makedirs(path, exist_ok=True)
    vocab_file = os.path.join(path, "vocab.txt")
    if not os.path.isfile(vocab_file):
        special = {PAD_ID: "[PAD]", UNK_ID: "[UNK]", CLS_ID: "[CLS]", SEP_ID: "[SEP]", MASK_ID: "[MASK]"}
        tmp = vocab_file + f".{os.getpid()}.tmp"
        with open(tmp, "w") as f:
            for i in range(vocab_size):
                f.write(special.get(i, f"[unused{i}]" if i < 1000 else f"w{i}") + "\n")
        with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
            json.dump({"tokenizer_class": "BertTokenizer", "do_lower_case": True}, f)
        os.replace(tmp, vocab_file)
    return path


def shipped_model_config(tokenizer: str, embed_dim: int = 768, enc_layers: int = 1, dec_layers: int = 3,
                         nhead: int = 8, feedforward: int = 2048, dropout: float = 0.3,
                         modal_shape=(512,), sce_loss_alpha: float = 0.5) -> dict:
    """The ``model`` block of configs/caption-task_baseline_modal_clip4clip_config.json:62-93
    (defaults = shipped values), restated so benchmarks do not need the reference tree."""
    return {
        "modal": ["CLIP4Clip"], "modal_shape": list(modal_shape), "tokenizer": tokenizer,
        "text_enc_type": "CLIP", "embed_dim": embed_dim, "dropout": dropout, "loss_beta": 0.5,
        "matching": {"enable_tem": False, "matching_loss": "CSL"}, "activation": "gelu",
        "video_encoder": {"layer": enc_layers, "nhead": nhead, "feedforward": feedforward,
                          "mme": {"temporal": "encoding", "modal_different": True, "do_norm": False,
                                  "aggregation": "avg"}, "aoa": False},
        "caption_decoder": {"layer": dec_layers, "nhead": nhead, "feedforward": feedforward,
                            "sce_loss_alpha": sce_loss_alpha},
        "pretrained_model": None,
    }


def synth_batch(B: int, T: int = 12, Din: int = 512, S1: int = 21, V: int = VOCAB_SIZE, seed: int = 1234,
                padded: bool = False):
    """feats fp32 [B,T,Din] ~ N(0,1); video pad mask all False; ids int64 [B,S1] with [CLS] first and
    [SEP] last (un-padded) or lengths ~ U{6..S1} with [SEP] then zeros (padded)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, Din, generator=g)
    tok = torch.randint(1000 if V > 2000 else 104, V, (B, S1), generator=g)
    tok[:, 0] = CLS_ID
    if padded:
        lens = torch.randint(min(6, S1 - 1), S1 + 1, (B,), generator=g)
        lens[0] = S1
        for b in range(B):
            tok[b, lens[b] - 1] = SEP_ID
            tok[b, lens[b]:] = PAD_ID
    else:
        tok[:, -1] = SEP_ID
    vm = torch.zeros(B, T, dtype=torch.bool)
    return x, vm, tok
