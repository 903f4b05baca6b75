"""Caption tokenisation (reference: model/CapPreprocessor.py).

Same contract -- ``(ids int64 [B, L], pad_mask bool [B, L])`` on ``device`` with [CLS] ... [SEP] and
``[PAD]`` = ``pad_id`` fill -- but the batch is assembled on the host and crosses PCIe once, instead of
one ``.to(device)`` per caption plus B row-assign kernels."""
from typing import List, Tuple

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
        enc = [self.tokenizer.encode(c) for c in captions]
        max_len = max(len(e) for e in enc)
        ids = torch.full((len(enc), max_len), self.pad_id, dtype=torch.long)
        for i, e in enumerate(enc):
            ids[i, :len(e)] = torch.tensor(e, dtype=torch.long)
        return ids

    def __call__(self, captions: List[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        ids = self.encode_host(captions)
        if torch.device(self.device).type == "cuda":
            ids = ids.pin_memory().to(self.device, non_blocking=True)
        else:
            ids = ids.to(self.device)
        return ids, ids == self.pad_id
