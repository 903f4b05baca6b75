"""CPU oracle for the Video-Captioning-Transformer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
(``video-captioning-transformer_b200/``); only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may use it, and only as the checker / CPU baseline.

This is a plain fp32 PyTorch *functional* restatement (explicit matmul /
softmax / layer-norm arithmetic over a ``state_dict``) of what the reference
computes through ``nn.TransformerEncoder`` / ``nn.TransformerDecoder``.  The
reference's arithmetic lives in un-vendored, un-pinned PyTorch (README only says
"torch 1.8.2+"); the semantics restated here are those of torch 2.11 run in this
image.  Parity status: **pinned** against the real reference modules imported
from ``/root/reference`` -- see ``oracle/make_golden.py`` (which generated
``tests/golden/*``) and ``tests/test_oracle.py`` (which re-checks the oracle
against those goldens everywhere, and against the live reference whenever
``/root/reference`` is mounted).  The reference itself ships no golden vectors
or tests for this path.

Citations are ``file:line`` into ``/root/reference`` unless they start with
``torch/``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
NEG_INF = float("-inf")


# --------------------------------------------------------------------------
# constant tables
# --------------------------------------------------------------------------
def sinusoid_table(maxlen: int, d: int) -> Tensor:
    """``[maxlen, d]`` sin/cos table.

    model/Embedding.py:13-17 (decoder positions, maxlen 5000) and
    model/MMEncoder.py:71-81 (temporal encoding, max_len 512) use the same
    formula: ``den = exp(-arange(0,d,2) * ln(10000)/d)``, even columns sin, odd
    columns cos.
    """
    den = torch.exp(-torch.arange(0, d, 2) * math.log(10000) / d)
    pos = torch.arange(0, maxlen).reshape(maxlen, 1)
    pe = torch.zeros((maxlen, d))
    pe[:, 0::2] = torch.sin(pos * den)
    pe[:, 1::2] = torch.cos(pos * den)
    return pe


def temporal_sinusoid_table(maxlen: int, d: int) -> Tensor:
    """``[maxlen, d]`` table of model/MMEncoder.py:71-81.  Same maths as ``sinusoid_table`` but the
    reference evaluates ``exp(arange(0,d,2).float() * -(ln(10000)/d))`` here (the scalar is divided
    first), which rounds differently in fp32 for some d -- restated literally so the buffer is
    bit-identical."""
    position = torch.arange(0, maxlen).float().unsqueeze(1)
    div_term = (torch.arange(0, d, 2).float() * -(math.log(10000.0) / d)).exp()
    pe = torch.zeros(maxlen, d)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def temporal_rows(pe: Tensor, T: int) -> Tensor:
    """``[T+1, d]`` temporal-encoding rows for ONE modality of length T.

    model/MMEncoder.py:89-104 (``separate=False`` branch): row 0 (the global
    token) is zero; row i+1 is ``pe[linspace(0, T-1, T)[i]]`` = ``pe[i]``.
    ``pe`` is the ``[1, 512, d]`` buffer or its ``[512, d]`` squeeze.
    """
    pe2 = pe.reshape(-1, pe.shape[-1])
    out = torch.zeros(T + 1, pe2.shape[-1], dtype=pe2.dtype)
    out[1:] = pe2[:T]
    return out


def causal_mask(S: int) -> Tensor:
    """Float ``[S,S]``: 0 on/below the diagonal, -inf above (utils.py:63-66)."""
    m = torch.full((S, S), NEG_INF)
    return torch.triu(m, diagonal=1)


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """LayerNorm over the last dim, biased variance, eps 1e-5
    (torch/nn/modules/transformer.py norm1/norm2/norm3, layer_norm_eps default)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """Exact-erf GELU: activation string "gelu" -> F.gelu(approximate='none')."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def mha(x_q: Tensor, x_kv: Tensor, in_w: Tensor, in_b: Tensor, out_w: Tensor, out_b: Tensor,
        nhead: int, add_mask: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """Multi-head attention, batch-first, returns (out [B,Lq,d], probs [B,h,Lq,Lk]).

    torch/nn/functional.py ``multi_head_attention_forward``: packed
    ``in_proj_weight [3d,d]`` in q,k,v order (``_in_projection_packed``; for
    cross-attention rows [0:d] act on the query input and rows [d:3d] on the
    memory), heads are contiguous ``dh`` slices, scale ``1/sqrt(dh)``, additive
    float mask (key padding merged with the attention mask), softmax over keys,
    then ``out_proj``.  Dropout is off (parity is defined at p=0 / eval).
    """
    B, Lq, d = x_q.shape
    Lk = x_kv.shape[1]
    dh = d // nhead
    q = x_q @ in_w[:d].T + in_b[:d]
    k = x_kv @ in_w[d:2 * d].T + in_b[d:2 * d]
    v = x_kv @ in_w[2 * d:].T + in_b[2 * d:]
    q = q.view(B, Lq, nhead, dh).transpose(1, 2)
    k = k.view(B, Lk, nhead, dh).transpose(1, 2)
    v = v.view(B, Lk, nhead, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(dh))
    if add_mask is not None:
        s = s + add_mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, d)
    return o @ out_w.T + out_b, p


def _ffn(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    h = gelu_erf(x @ sd[pre + "linear1.weight"].T + sd[pre + "linear1.bias"])
    return h @ sd[pre + "linear2.weight"].T + sd[pre + "linear2.bias"]


def _count_layers(sd: Dict[str, Tensor], prefix: str) -> int:
    n = 0
    while (prefix + f"{n}.linear1.weight") in sd:
        n += 1
    return n


# --------------------------------------------------------------------------
# encoder  (model/MMEncoder.py:244-276, single modality, temporal "encoding",
#           aggregation "avg", do_norm False -- the branch the shipped JSON selects)
# --------------------------------------------------------------------------
def encoder_forward(sd: Dict[str, Tensor], feats: Tensor, pad_mask: Optional[Tensor], nhead: int,
                    prefix: str = "video_encoder.", eval_fastpath: bool = False) -> Tensor:
    """feats [B,T,Din] fp32, pad_mask bool [B,T] (True = ignore) or None -> memory [B,T+1,d].

    eval_fastpath: what ``val_epoch`` (train.py:151-168) and ``eval.py:140`` actually execute -- ``eval()`` +
    ``no_grad`` + a key-padding mask make ``nn.TransformerEncoder`` take torch's nested-tensor fast path
    (torch/nn/modules/transformer.py:452-548), which zero-fills the padded positions before the final LayerNorm:
    memory rows of padded frames become ``LayerNorm(0) = norm.bias``; valid rows are unchanged."""
    B, T, _ = feats.shape
    u = feats @ sd[prefix + "unify.0.weight"].T + sd[prefix + "unify.0.bias"]          # :246
    g = u.mean(dim=1, keepdim=True)            # :248-250, avg over ALL T incl. padded frames (Q4)
    x = torch.cat([g, u], dim=1)               # [B, M, d]
    x = x + temporal_rows(sd[prefix + "temp_emb.pe"], T).unsqueeze(0)                   # :262-271
    add_mask = None
    if pad_mask is not None:                   # :252-260: prepend False for the global token
        full = torch.cat([torch.zeros(B, 1, dtype=torch.bool), pad_mask], dim=1)
        add_mask = torch.zeros(B, 1, 1, T + 1).masked_fill(full[:, None, None, :], NEG_INF)
    lp = prefix + "transformer_encoder.layers."
    for i in range(_count_layers(sd, lp)):     # post-norm, torch/nn/modules/transformer.py:946-982
        p = lp + f"{i}."
        a, _ = mha(x, x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                   sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead, add_mask)
        x = layer_norm(x + a, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = layer_norm(x + _ffn(x, sd, p), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    if eval_fastpath and pad_mask is not None:
        x = x.masked_fill(full[:, :, None], 0.0)
    return layer_norm(x, sd[prefix + "transformer_encoder.norm.weight"], sd[prefix + "transformer_encoder.norm.bias"])


# --------------------------------------------------------------------------
# decoder  (model/CapDecoder.py:34-79, torch/nn/modules/transformer.py:1131-1199)
# --------------------------------------------------------------------------
def decoder_hidden(sd: Dict[str, Tensor], memory: Tensor, tgt_in: Tensor, pad_in: Optional[Tensor], nhead: int,
                   prefix: str = "cap_decoder.", return_cross_probs: bool = False):
    """Teacher-forced decoder stack: token ids [B,S] -> hidden [B,S,d] (after the final LN).

    Embedding has no sqrt(d) scaling (Q7); self-attention gets causal + key
    padding; cross-attention is NEVER masked (Q3, model/CapDecoder.py:49-52).
    """
    B, S = tgt_in.shape
    x = sd[prefix + "tgt_to_emb.weight"][tgt_in] + sd[prefix + "positional_encoding.pos_embedding"][:S]
    add_mask = causal_mask(S)[None, None]
    if pad_in is not None:
        add_mask = add_mask + torch.zeros(B, 1, 1, S).masked_fill(pad_in[:, None, None, :], NEG_INF)
    lp = prefix + "decoder.layers."
    cross = []
    for i in range(_count_layers(sd, lp)):
        p = lp + f"{i}."
        a, _ = mha(x, x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                   sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead, add_mask)
        x = layer_norm(x + a, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        c, pc = mha(x, memory, sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"],
                    sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"], nhead, None)
        cross.append(pc.mean(dim=1))           # head-averaged, what need_weights=True returns
        x = layer_norm(x + c, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        x = layer_norm(x + _ffn(x, sd, p), sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    h = layer_norm(x, sd[prefix + "decoder.norm.weight"], sd[prefix + "decoder.norm.bias"])
    return (h, cross) if return_cross_probs else h


def generator(sd: Dict[str, Tensor], h: Tensor, prefix: str = "cap_decoder.") -> Tensor:
    """model/CapDecoder.py:25,55."""
    return h @ sd[prefix + "generator.weight"].T + sd[prefix + "generator.bias"]


# --------------------------------------------------------------------------
# loss  (model/loss.py:69-92), closed form of SURVEY Q9
# --------------------------------------------------------------------------
RCE_A = -math.log(1e-4)


def sce_loss(logits: Tensor, labels: Tensor, alpha: float, beta: float, pad_id: int = 0) -> Tensor:
    """alpha * CE(mean over non-pad rows) + beta * mean over ALL rows of
    A * sum_{c != y} clamp(softmax(z)_c, 1e-7, 1)   with A = -ln(1e-4).

    If alpha == 1.0 the reference uses plain CrossEntropyLoss(ignore_index=pad)
    instead (model/CapDecoder.py:28-32).
    """
    lse = torch.logsumexp(logits, dim=1)
    zy = logits.gather(1, labels[:, None]).squeeze(1)
    valid = labels != pad_id
    ce = ((lse - zy) * valid).sum() / valid.sum()
    if alpha == 1.0:
        return ce
    p = torch.clamp(torch.softmax(logits, dim=1), min=1e-7, max=1.0)
    py = p.gather(1, labels[:, None]).squeeze(1)
    rce = RCE_A * (p.sum(dim=1) - py)
    return alpha * ce + beta * rce.mean()


# --------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------
def caption_forward(sd: Dict[str, Tensor], feats: Tensor, vid_pad: Optional[Tensor], ids: Tensor,
                    enc_nhead: int, dec_nhead: int, alpha: float, pad_id: int = 0, eval_fastpath: bool = False
                    ) -> Tuple[Tensor, Tensor, Tensor]:
    """model/MMT4Caption.py:114-121 with pre-tokenised ids [B,S+1].

    Returns (memory, logits [B,S,V], loss).  tgt_in = ids[:, :-1], tgt_out =
    ids[:, 1:], pad = (ids == pad_id)[:, :-1]  (model/CapDecoder.py:43-45).
    """
    memory = encoder_forward(sd, feats, vid_pad, enc_nhead, eval_fastpath=eval_fastpath)
    tgt_in, tgt_out = ids[:, :-1], ids[:, 1:]
    pad_in = (ids == pad_id)[:, :-1]
    h = decoder_hidden(sd, memory, tgt_in, pad_in, dec_nhead)
    logits = generator(sd, h)
    loss = sce_loss(logits.reshape(-1, logits.shape[-1]), tgt_out.reshape(-1), alpha, 1.0 - alpha, pad_id)
    return memory, logits, loss


def caption_grads(sd: Dict[str, Tensor], feats: Tensor, vid_pad: Optional[Tensor], ids: Tensor,
                  enc_nhead: int, dec_nhead: int, alpha: float, pad_id: int = 0,
                  frozen_prefixes: Tuple[str, ...] = ("matching.",)) -> Tuple[Tensor, Dict[str, Tensor]]:
    """loss and d(loss)/d(param) for every float parameter (autograd over the restatement)."""
    buffers = ("pos_embedding", "temp_emb.pe")
    leaf = {}
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(buffers) and not k.startswith(frozen_prefixes):
            leaf[k] = v.detach().clone().requires_grad_(True)
        else:
            leaf[k] = v
    _, _, loss = caption_forward(leaf, feats, vid_pad, ids, enc_nhead, dec_nhead, alpha, pad_id)
    names = [k for k, v in leaf.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
    out = {}
    for k, g in zip(names, grads):
        out[k] = torch.zeros_like(leaf[k]) if g is None else g
    # nn.Embedding(padding_idx=pad_id): the pad row never receives gradient (Q7)
    out["cap_decoder.tgt_to_emb.weight"][pad_id].zero_()
    return loss.detach(), out


def decode_word(sd: Dict[str, Tensor], memory: Tensor, ys: Tensor, dec_nhead: int) -> Tensor:
    """model/CapDecoder.py:62-79: full re-run over all tokens, logits of the last position."""
    h = decoder_hidden(sd, memory, ys, None, dec_nhead)
    return generator(sd, h[:, -1])


def greedy_decode_ids(sd: Dict[str, Tensor], feats: Tensor, vid_pad: Optional[Tensor], enc_nhead: int,
                      dec_nhead: int, max_len: int = 30, start_id: int = 101, end_id: int = 102,
                      eval_fastpath: bool = False, return_margins: bool = False):
    """model/MMT4Caption.py:146-172: returns ys [B, <=max_len] (incl. the start token).

    All rows keep generating until every row has produced ``end_id`` at least
    once (Q11); argmax ties resolve to the lowest index (torch.max).
    """
    B = feats.shape[0]
    memory = encoder_forward(sd, feats, vid_pad, enc_nhead, eval_fastpath=eval_fastpath)
    ys = torch.full((B, 1), start_id, dtype=torch.long)
    ended = torch.zeros(B, dtype=torch.bool)
    margins = []
    for _ in range(max_len - 1):
        lg = decode_word(sd, memory, ys, dec_nhead)
        nxt = lg.argmax(dim=1)
        if return_margins:
            t2 = lg.topk(2, dim=1).values
            margins.append(t2[:, 0] - t2[:, 1])
        ys = torch.cat([ys, nxt[:, None]], dim=1)
        ended |= nxt == end_id
        if bool(ended.all()):
            break
    return (ys, torch.stack(margins, 1)) if return_margins else ys


def cut_caption_ids(row: List[int], end_id: int = 102) -> List[int]:
    """model/MMT4Caption.py:175-181: drop [CLS]; cut at the first [SEP]; when there is
    none, ``end_count = -1`` so ``ids[1:-1]`` drops the last token (Q11)."""
    end_count = -1
    for i, t in enumerate(row):
        if t == end_id:
            end_count = i
            break
    return row[1:end_count]


# --------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults used by train.py:22-31: amsgrad False, wd 0)
# --------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, wd: float = 0.0):
    """One Adam update, returns (p, m, v).  ``step`` is 1-based."""
    if wd != 0.0:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v
