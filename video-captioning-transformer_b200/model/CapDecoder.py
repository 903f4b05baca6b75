"""Caption decoder (reference: model/CapDecoder.py:11-79).

Same constructor, attributes and state_dict keys: the ``nn.TransformerDecoder`` container, ``generator``,
``tgt_to_emb`` and ``positional_encoding`` are instantiated exactly as the reference does (identical RNG
consumption) and serve as PARAMETER HOLDERS -- their ``forward`` is never called on the hot
path.  ``forward`` / ``decode_word`` run the fused B200 kernels through ``vct.engine.CaptionEngine``:
embedding+positional add, per layer {packed in-projection + causal/key-padding attention, out-projection,
residual+dropout+LayerNorm, cross-attention over the encoder memory (never masked, Q3), FFN with fused
bias+GELU+dropout}, final LayerNorm, generator GEMM and the single-pass SCE loss kernel."""
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .Embedding import PositionalEmbedding
from .loss import SCELoss
from ._engine import build_engine, param_device


def generate_square_subsequent_mask(sz: int) -> Tensor:
    """Float [sz, sz] causal mask: 0 on/below the diagonal, -inf above.  The reference imports this from
    its top-level utils.py (:63-66); the kernels implement the mask implicitly (key j > query i), this
    helper exists for API compatibility."""
    return torch.triu(torch.full((sz, sz), float('-inf')), diagonal=1)


class CapDecoder(nn.Module):
    def __init__(self, num_layers, embed_dim, nhead, dim_feedforward, dropout,
                 vocab_size, pad_id, sce_loss_alpha: float, custom_decoder_type: Optional[str] = None,
                 activation='gelu', device=torch.device('cuda')):
        super().__init__()
        self.device = device
        if custom_decoder_type is not None:
            raise NotImplementedError("caption_decoder.layer_type (VisTransformerDecoder) is outside the hot path: "
                                      "the shipped configs do not set it (SURVEY section 2.1 #3b)")
        if activation != 'gelu':
            raise NotImplementedError("only activation='gelu' (the shipped configs' value) has a fused kernel")
        decoder_layer = nn.TransformerDecoderLayer(embed_dim, nhead, dim_feedforward, dropout,
                                                   activation=activation, batch_first=True)
        self.decoder = nn.TransformerDecoder(decoder_layer, num_layers, nn.LayerNorm(embed_dim))
        self.generator = nn.Linear(embed_dim, vocab_size)
        self.tgt_to_emb = nn.Embedding(vocab_size, embed_dim, padding_idx=pad_id)
        self.positional_encoding = PositionalEmbedding(embed_dim, dropout=dropout, maxlen=5000)
        self.pad_id = pad_id
        self.sce_loss_alpha = float(sce_loss_alpha)
        self.dropout_p = float(dropout)
        if sce_loss_alpha == 1.0:
            self.loss_fn = nn.CrossEntropyLoss(ignore_index=pad_id)
        else:
            self.loss_fn = SCELoss(sce_loss_alpha, 1 - sce_loss_alpha, ignore_index=pad_id, num_classes=vocab_size,
                                   device=device)
        self._vct_engine = None

    # ---- engine plumbing ---------------------------------------------------------------------------
    def _engine(self):
        owner = self.__dict__.get("_vct_owner")
        if owner is not None and owner() is not None:
            return owner()._engine()            # part of an MMT4Caption: one joint engine / arena
        eng = self._vct_engine
        if eng is None or not eng.arena.is_current():
            eng = build_engine(None, self, param_device(self), precision=self.__dict__.get("vct_precision"), gemm_impl=self.__dict__.get("vct_gemm"))
            object.__setattr__(self, "_vct_engine", eng)
        return eng

    def _uses_patched_layers(self) -> bool:
        """predict_video.py:126-130 rebinds ``layer.forward`` on every decoder layer."""
        return any('forward' in layer.__dict__ for layer in self.decoder.layers)

    # ---- reference API ------------------------------------------------------------------------------
    def forward(self, memories: Tensor, tgt: Tensor, tgt_padding_mask: Optional[Tensor], _want_logits: bool = True):
        """memories [B,M,E], tgt ids [B,S+1], tgt_padding_mask [B,S+1] (True = pad) -> (logits [B,S,V], loss).
        As in the reference (model/CapDecoder.py:43-52) the decoder's key-padding mask is ``tgt_padding_mask[:, :-1]``
        -- the mask that is PASSED, not one recomputed from the ids -- while the loss ignores positions whose target
        id is ``pad_id`` (model/CapDecoder.py:28-32).  ``None`` means ``tgt == pad_id``.  Both outputs are freshly
        allocated."""
        from vct.functional import DecoderFn
        eng = self._engine()
        params = [p for _, p in self.named_parameters()]
        tok_pad = None
        if tgt_padding_mask is not None:
            if tgt_padding_mask.shape != tgt.shape:
                raise ValueError(f"tgt_padding_mask {tuple(tgt_padding_mask.shape)} must match tgt {tuple(tgt.shape)}")
            tok_pad = tgt_padding_mask[:, :-1].to(torch.bool)
        return DecoderFn.apply(eng, self, memories, tgt, tok_pad, bool(_want_logits), *params)

    def decode_word(self, memories: Tensor, tgt: Tensor, tgt_padding_mask: Optional[Tensor]):
        """Next-word logits [B,V] for the prefix ``tgt`` [B,t] (model/CapDecoder.py:62-79).  Runs the
        teacher-forced decoder kernels over the prefix (the reference recomputes every position too);
        ``MMT4Caption.greedy_decode`` uses the K/V-cached incremental plan instead."""
        if tgt_padding_mask is not None:
            raise NotImplementedError("decode_word with a padding mask is not used by the reference "
                                      "(model/MMT4Caption.py:164 passes None)")
        eng = self._engine()
        B, t = tgt.shape
        M = memories.shape[1]
        with torch.no_grad():
            ws = eng.workspace(B, M - 1, t, False)
            eng.check_arena()
            eng.refresh_shadow()
            if memories.data_ptr() != ws.mem.data_ptr():
                ws.mem.copy_(memories.reshape(ws.mem.shape))
                if ws.mem_c is not ws.mem:
                    ws.mem_c.copy_(ws.mem)
            ws.ids[:, :t].copy_(tgt)
            ws.ids[:, t] = self.pad_id
            ws.tok_pad.zero_()
            eng.run(eng.plan_forward(ws, fused_grad=False, part="dec", with_loss=False))
            V = eng.dims.V
            return ws.logits.view(B, t, ws.Vp)[:, -1, :V].clone()
