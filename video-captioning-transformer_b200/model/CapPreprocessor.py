"""Caption tokenisation (reference: model/CapPreprocessor.py).

Same contract -- ``(ids int64 [B, L], pad_mask bool [B, L])`` on ``device`` with [CLS] ... [SEP] and
``[PAD]`` = ``pad_id`` fill -- but the batch is assembled on the host and crosses PCIe once, instead of
one ``.to(device)`` per caption plus B row-assign kernels."""
from typing import List, Tuple

import numpy as np
import torch
from transformers import AutoTokenizer


class CapPreprocessor:
    def __init__(self, tokenizer_type, device=torch.device('cuda')):
        self.tokenizer_type = tokenizer_type
        self.device = device
        self.tokenizer = AutoTokenizer.from_pretrained(tokenizer_type)
        self.pad_id = self.tokenizer.convert_tokens_to_ids("[PAD]")
        self.start_id = self.tokenizer.convert_tokens_to_ids("[CLS]")
        self.end_id = self.tokenizer.convert_tokens_to_ids("[SEP]")

    def encode_host(self, captions: List[str]) -> torch.Tensor:
        """[B, L] int64 on the host.  ONE batched tokenizer call for the captions not seen before (the reference encodes
        caption by caption, model/CapPreprocessor.py:24-28; a training set repeats every caption each epoch, so the ids are
        memoised per string), rows assembled in one numpy buffer."""
        cache = self.__dict__.setdefault("_ids_cache", {})
        new = [c for c in dict.fromkeys(captions) if c not in cache]
        if new:
            try:
                enc = self.tokenizer(new, add_special_tokens=True, padding=False, truncation=False)["input_ids"]
            except Exception:                                    # a tokenizer without a batch entry point
                enc = [self.tokenizer.encode(c) for c in new]
            if len(cache) + len(new) > 2_000_000:
                cache.clear()
            for c, e in zip(new, enc):
                cache[c] = np.asarray(e, dtype=np.int64)
        rows = [cache[c] for c in captions]
        max_len = max(len(e) for e in rows)
        ids = np.full((len(rows), max_len), self.pad_id, dtype=np.int64)
        for i, e in enumerate(rows):
            ids[i, :len(e)] = e
        return torch.from_numpy(ids)

    def __call__(self, captions: List[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        ids = self.encode_host(captions)
        if torch.device(self.device).type == "cuda":
            ids = ids.pin_memory().to(self.device, non_blocking=True)
        else:
            ids = ids.to(self.device)
        return ids, ids == self.pad_id
