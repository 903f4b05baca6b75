"""Frozen text tower for the "match"/"cross" tasks (reference: model/TextEncoder.py).  Not on the
caption hot path; a plain object (no parameters in state_dict / DDP), constructed lazily so that the
caption task works offline when neither ``clip`` nor BERT weights are available."""
from typing import List

import torch
from torch import Tensor


class TextEncoder:
    def __init__(self, text_enc_type: str, device=torch.device('cuda')):
        self.text_enc_type = text_enc_type
        self.device = device
        self.text_enc = None
        if "CLIP" in text_enc_type:
            self.dim = 512
        elif "bert" in text_enc_type:
            self.dim = 768
        else:
            raise ValueError

    def _load(self):
        if self.text_enc is not None:
            return
        if "CLIP" in self.text_enc_type:
            import clip
            self.text_enc, _ = clip.load("ViT-B/32", device=self.device)
            self.text_enc.eval()
        else:
            from transformers import AutoTokenizer, BertModel
            self.tokenizer = AutoTokenizer.from_pretrained("./data/tk")
            self.text_enc = BertModel.from_pretrained(self.text_enc_type).to(self.device)

    def __call__(self, captions: List[str]) -> Tensor:
        self._load()
        if "CLIP" in self.text_enc_type:
            import clip
            tokens = clip.tokenize(captions).to(self.device)
            return self.text_enc.encode_text(tokens).to(torch.float32).detach()
        pad_id = self.tokenizer.convert_tokens_to_ids("[PAD]")
        enc = [self.tokenizer.encode(c) for c in captions]
        max_len = max(len(e) for e in enc)
        text_ts = torch.full((len(enc), max_len), pad_id, dtype=torch.long)
        for i, e in enumerate(enc):
            text_ts[i, :len(e)] = torch.tensor(e)
        text_ts = text_ts.to(self.device)
        return self.text_enc(text_ts, text_ts == pad_id).last_hidden_state[:, 0]
