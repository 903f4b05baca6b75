"""Video encoder (reference: model/MMEncoder.py).

``MultiModalEncoder`` keeps the reference constructor, attribute names and state_dict keys
(``unify.0.*``, ``temp_emb.pe``, ``transformer_encoder.*``) and runs the branch the shipped configs select
(one modality, temporal "encoding", aggregation "avg", do_norm false) on the fused
B200 kernels: frame staging + unify GEMM whose epilogue adds bias and the temporal table (the global
"avg" token is produced by the same GEMM from the mean frame, unify being linear), then per layer packed
in-projection + key-padding attention, out-projection, residual+dropout+LayerNorm, FFN, and the final
LayerNorm.  The other encoder variants of the reference (learned temporal/modal embeddings, GRU/max
aggregation, SimpleSepEncoder, HMMEncoder) are not selected by any shipped config and are outside the hot
path; selecting them raises instead of silently running an un-fused path."""
import math
from typing import List, Optional

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn import ModuleList

from ._engine import build_engine, param_device


class TemporalEncoding(nn.Module):
    """Holds the constant ``pe [1, max_len, d]`` buffer (model/MMEncoder.py:51-81).  The per-forward host
    loop of the reference (:89-104) is replaced by a table cached per T inside the engine."""

    def __init__(self, d_model=512, max_len=512, separate=False, device=torch.device("cuda")):
        super().__init__()
        self.d_model = d_model
        self.device = device
        self.separate = separate
        # same operation order as the reference so the buffer is bit-identical to its checkpoints
        position = torch.arange(0, max_len).float().unsqueeze(1)
        div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
        pe = torch.zeros(max_len, d_model).float()
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0).to(device))

    def rows(self, T: int) -> Tensor:
        """[T+1, d]: zeros for the global token, then pe[0..T-1]."""
        out = torch.zeros(T + 1, self.d_model, dtype=self.pe.dtype, device=self.pe.device)
        out[1:] = self.pe[0, :T]
        return out


class GlobalAggregation(nn.Module):
    """Parameter-free holder for the "avg" aggregation (model/MMEncoder.py:173-201); the mean over ALL T
    frames (padded ones included) is computed by the vct_prep_frames kernel."""

    def __init__(self, method: str = "max", d_model: Optional[int] = None, device=torch.device("cuda")):
        super().__init__()
        if method != "avg":
            raise NotImplementedError(f"aggregation '{method}' is outside the hot path (shipped configs use 'avg')")
        self.method = method
        self.device = device
        self.agg = nn.AdaptiveAvgPool1d(1)


class MultiModalEncoder(nn.Module):
    def __init__(self, d_feats: List[int], d_model: int, nhead: int,
                 dim_feedforward: int = 2048, num_encoder_layers: int = 4,
                 dropout: float = 0.1, activation: str = "gelu", global_type: str = "avg",
                 modal_different: bool = True, temporal_type: str = "embedding", do_norm: bool = False,
                 device=torch.device("cuda")):
        super().__init__()
        self.device = device
        self.num_modal = len(d_feats)
        self.do_norm = do_norm
        if self.num_modal != 1:
            raise NotImplementedError("multi-modal input (ModalEmbedding) is outside the hot path: the shipped "
                                      "configs use a single CLIP4Clip modality")
        if temporal_type == "embedding":
            raise NotImplementedError("temporal 'embedding' (learned) is outside the hot path: the shipped configs "
                                      "use temporal 'encoding'")
        if do_norm:
            raise NotImplementedError("do_norm=True is outside the hot path (shipped configs: false)")
        if activation != "gelu":
            raise NotImplementedError("only activation='gelu' (the shipped configs' value) has a fused kernel")
        self.unify = ModuleList([nn.Linear(d_feat, d_model) for d_feat in d_feats])
        self.global_agg = GlobalAggregation(global_type, d_model=d_model, device=device)
        self.temp_emb = TemporalEncoding(d_model, device=device)
        encoder_layer = nn.TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout,
                                                   activation=activation, batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(encoder_layer, num_encoder_layers, nn.LayerNorm(d_model))
        self.dropout_p = float(dropout)
        self._vct_engine = None
        self._vct_S_hint = 1

    def _engine(self):
        owner = self.__dict__.get("_vct_owner")
        if owner is not None and owner() is not None:
            return owner()._engine()            # part of an MMT4Caption: one joint engine / arena
        eng = self._vct_engine
        if eng is None or not eng.arena.is_current():
            eng = build_engine(self, None, param_device(self), precision=self.__dict__.get("vct_precision"), gemm_impl=self.__dict__.get("vct_gemm"))
            object.__setattr__(self, "_vct_engine", eng)
        return eng

    def forward(self, srcs: List[Tensor], src_padding_masks: Optional[List[Tensor]]):
        """srcs: [Tensor[B,T,Din]] (one modality); masks: [Bool[B,T]] (True = padded) or None
        -> (memory [B,T+1,E], global_masks [B,T+1] | None, memory[:, 0])   (model/MMEncoder.py:244-276).
        Padded frames: in training (and whenever gradients are enabled) their memory rows hold the values the layers
        compute, as in the reference; under eval() + no_grad with masks -- val_epoch / eval.py -- the reference's
        nn.TransformerEncoder takes torch's nested-tensor fast path and those rows become norm.bias, which
        is reproduced here (tests/golden "evalfast")."""
        from vct.functional import EncoderFn
        if len(srcs) != 1:
            raise NotImplementedError("single modality only (see class docstring)")
        eng = self._engine()
        feats = srcs[0]
        mask = src_padding_masks[0] if src_padding_masks is not None else None
        params = [p for _, p in self.named_parameters()]
        fast = (not self.training) and (not torch.is_grad_enabled()) and mask is not None
        memory = EncoderFn.apply(eng, self, feats, mask, int(self._vct_S_hint), fast, *params)
        global_masks = None
        if mask is not None:
            global_masks = torch.cat([torch.zeros(mask.shape[0], 1, dtype=torch.bool, device=mask.device), mask], dim=1)
        return memory, global_masks, memory[:, 0]


class SimpleSepEncoder(nn.Module):
    """model/MMEncoder.py:280-310 -- ``video_encoder.type == "simple"``; not selected by shipped configs."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("video_encoder.type 'simple' is outside the hot path (SURVEY section 2.1 #2d)")


class HMMEncoder(nn.Module):
    """model/MMEncoder.py:314-402 -- ``video_encoder.type == "hmme"``; not selected by shipped configs."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("video_encoder.type 'hmme' is outside the hot path (SURVEY section 2.1 #2d)")
