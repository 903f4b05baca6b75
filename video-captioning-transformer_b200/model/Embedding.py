"""Decoder positional table (reference: model/Embedding.py:7-25).

Only the constant buffer matters on the hot path: ``CapDecoder`` fuses
``dropout(E[ids] + pos_embedding[:S])`` into the ``vct_embed_fwd`` kernel, so this module's
``forward`` is API surface for callers that hold an embedding tensor already."""
import math

import torch
import torch.nn as nn
from torch import Tensor


def sinusoid_table(maxlen: int, emb_size: int) -> Tensor:
    """[maxlen, emb_size]; even columns sin, odd columns cos of pos * exp(-2i ln(10000)/emb_size).
    Evaluated in the reference's operation order (model/Embedding.py:13-17) so the registered buffer is
    bit-identical to a reference checkpoint's ``positional_encoding.pos_embedding``."""
    den = torch.exp(- torch.arange(0, emb_size, 2) * math.log(10000) / emb_size)
    pos = torch.arange(0, maxlen).reshape(maxlen, 1)
    table = torch.zeros((maxlen, emb_size))
    table[:, 0::2] = torch.sin(pos * den)
    table[:, 1::2] = torch.cos(pos * den)
    return table


class PositionalEmbedding(nn.Module):
    def __init__(self, emb_size: int, dropout: float, maxlen: int = 5000):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.register_buffer('pos_embedding', sinusoid_table(maxlen, emb_size))

    def forward(self, token_embedding: Tensor):
        # API surface only (not on the fused hot path, see module docstring)
        return self.dropout(token_embedding + self.pos_embedding[:token_embedding.shape[1], :])
