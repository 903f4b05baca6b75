"""torch.autograd bridges: they let the reference's own training loop
(``loss = model(...); loss.backward(); optimizer.step()``, train.py:123-126) and
DistributedDataParallel drive the B200 kernels.  Forward and backward are launch plans of
``CaptionEngine``; these classes only move gradients in and out of the arena.

Contract of this path (the reference returns freshly allocated tensors):
  * outputs handed to the caller (memory, logits, loss) are COPIES of the workspace buffers, so a later forward of the
    same shape cannot change tensors the caller still holds;
  * activations saved for backward stay in the per-shape workspace: running another forward of the same shape
    before ``backward()`` invalidates them, which is detected and raised (never silently wrong gradients);
  * every training forward draws a fresh dropout step (``engine.next_rng_step``) and its backward re-installs that
    step first, so forward and backward masks are identical and consecutive calls differ."""
from __future__ import annotations

from typing import Optional

import torch

from .engine import CaptionEngine


def _detach_aliased_grads(engine: CaptionEngine, prefix: str, params) -> None:
    """Gradient accumulation (a second ``backward()`` without ``zero_grad(set_to_none=True)``): a ``p.grad`` that
    AccumulateGrad stole from a previous backward is a VIEW of the gradient arena, which the plan about to run
    overwrites.  Move such gradients to their own storage first so that autograd adds old + new correctly."""
    a = engine.arena
    base, end = a.grad.data_ptr(), a.grad.data_ptr() + 4 * a.numel
    for _name, p in params:
        g = p.grad
        if g is not None and base <= g.data_ptr() < end:
            p.grad = g.clone()


def _grads_for(engine: CaptionEngine, prefix: str, params):
    """Arena gradient views for the given parameters: zero-copy when ``p.grad`` is unset (what
    ``optimizer.zero_grad()`` leaves -- AccumulateGrad then adopts the view), a clone when autograd is going to
    accumulate into an existing ``p.grad``."""
    out = []
    a = engine.arena
    for name, p in params:
        if not p.requires_grad:
            out.append(None)
            continue
        g = a.grad_view(prefix + name)
        out.append(g if p.grad is None else g.clone())
    return out


def _stale(what: str):
    return RuntimeError(f"vct: the {what} workspace was reused by another forward of the same shape before this "
                        f"backward ran; call backward() before the next forward of that shape (activations are saved "
                        f"in a per-shape workspace, not per call)")


class EncoderFn(torch.autograd.Function):
    """MultiModalEncoder.forward (model/MMEncoder.py:244-276) -> memory [B, T+1, d]."""

    @staticmethod
    def forward(ctx, engine: CaptionEngine, module, feats, vid_pad, S_hint: int, zero_pad: bool, *params):
        B, T, _ = feats.shape
        training = module.training
        ws = engine.workspace(B, T, S_hint, training)
        engine.check_arena()
        engine.refresh_shadow()
        ctx.rng_step = engine.next_rng_step() if training else None
        engine.stage_inputs(ws, feats, vid_pad, None)
        engine.run(engine.plan_encode(ws, zero_pad=bool(zero_pad) and not training))
        ws.enc_version += 1
        ctx.engine, ctx.ws, ctx.module, ctx.version = engine, ws, module, ws.enc_version
        out = ws.mem.view(B, T + 1, engine.dims.d).clone()
        # lets DecoderFn recognise "this is the memory my workspace already holds" without comparing contents
        ws.mem_token = (out.data_ptr(), out._version, ws.enc_version)
        return out

    @staticmethod
    def backward(ctx, dmem):
        engine, ws = ctx.engine, ctx.ws
        if not ws.training:
            raise RuntimeError("vct: backward through an eval-mode encoder forward is not supported "
                               "(call model.train(); dropout p can be 0)")
        if ws.enc_version != ctx.version:
            raise _stale("encoder")
        named = list(ctx.module.named_parameters())
        _detach_aliased_grads(engine, "video_encoder.", named)
        engine.set_rng_step(ctx.rng_step)
        if dmem.data_ptr() != ws.g_mem.data_ptr():
            ws.g_mem.copy_(dmem.reshape(ws.g_mem.shape))
        engine.run(engine.plan_backward(ws, sce_first=False, part="enc"))
        grads = _grads_for(engine, "video_encoder.", named)
        return (None, None, None, None, None, None, *grads)


class DecoderFn(torch.autograd.Function):
    """CapDecoder.forward (model/CapDecoder.py:34-60) -> (logits [B,S,V], loss)."""

    @staticmethod
    def forward(ctx, engine: CaptionEngine, module, memory, ids, tok_pad, want_logits: bool, *params):
        B, M, d = memory.shape
        S = ids.shape[1] - 1
        training = module.training
        ws = engine.workspace(B, M - 1, S, training)
        engine.check_arena()
        engine.refresh_shadow()
        ctx.rng_step = engine.next_rng_step() if training else None
        tok = ws.mem_token
        if tok is None or tok != (memory.data_ptr(), memory._version, ws.enc_version):
            # memory produced elsewhere (stand-alone use of the decoder, or edited by the caller): stage it
            ws.mem.copy_(memory.reshape(ws.mem.shape))
            if ws.mem_c is not ws.mem:
                ws.mem_c.copy_(ws.mem)
            ws.mem_token = None
        engine.stage_ids(ws, ids, tok_pad)
        engine.run(engine.plan_forward(ws, fused_grad=False, part="dec"))
        ws.dec_version += 1
        ctx.engine, ctx.ws, ctx.module, ctx.version = engine, ws, module, ws.dec_version
        V = engine.dims.V
        if want_logits:
            logits = ws.logits.view(B, S, ws.Vp)[:, :, :V].to(torch.float32, copy=True)
        else:
            logits = ws.logits.new_empty(0)         # caption_forward discards the logits (model/MMT4Caption.py:120-121)
        loss = ws.loss[0].clone()
        ctx.mark_non_differentiable(logits)
        return logits, loss

    @staticmethod
    def backward(ctx, _dlogits, dloss):
        engine, ws = ctx.engine, ctx.ws
        if not ws.training:
            raise RuntimeError("vct: backward through an eval-mode decoder forward is not supported "
                               "(call model.train(); dropout p can be 0)")
        if ws.dec_version != ctx.version:
            raise _stale("decoder")
        named = list(ctx.module.named_parameters())
        _detach_aliased_grads(engine, "cap_decoder.", named)
        engine.set_rng_step(ctx.rng_step)
        engine.upstream.copy_(dloss.reshape(1))
        engine.zero_scatter_grads()
        engine.run(engine.plan_backward(ws, sce_first=True, part="dec"))
        grads = _grads_for(engine, "cap_decoder.", named)
        # (a view: the encoder's backward consumes it at once; nothing else of this path keeps it)
        dmem = ws.g_mem.view(ws.B, ws.M, engine.dims.d) if ctx.needs_input_grad[2] else None
        return (None, None, dmem, None, None, None, *grads)


def sce_loss(engine_lib, logits: torch.Tensor, labels: torch.Tensor, alpha: float, beta: float, pad_id: int):
    """Stand-alone SCELoss.forward(pred [N,V], labels [N]) (model/loss.py:78-92) on the vct_sce kernel."""
    return _SceFn.apply(logits, labels, alpha, beta, pad_id)


class _SceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, alpha, beta, pad_id):
        from . import lib as L
        lib = L.load()
        if not logits.is_cuda:
            raise RuntimeError("vct_b200 has no CPU path: SCELoss needs CUDA tensors")
        N, V = logits.shape
        Vp = (V + 7) // 8 * 8
        z = torch.empty((N, Vp), dtype=torch.float32, device=logits.device)
        z[:, :V] = logits
        lab = torch.empty(N + 1, dtype=torch.int64, device=logits.device)
        lab[1:] = labels
        loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
        parts = torch.empty((N, 2), dtype=torch.float32, device=logits.device)
        counter = torch.zeros(1, dtype=torch.int32, device=logits.device)
        stream = torch.cuda.current_stream(logits.device).cuda_stream
        # label of row i is ids[i*1 + 0 + 1] with B = N, S = 1, ids_ld = 1
        L.check(lib.vct_sce(z.data_ptr(), Vp, lab.data_ptr(), 1, N, 1, V, float(alpha), float(beta), int(pad_id),
                            loss.data_ptr(), parts.data_ptr(), counter.data_ptr(), None, L.F32, Vp, None, stream), "vct_sce")
        ctx.save_for_backward(z, lab)
        ctx.meta = (N, V, Vp, float(alpha), float(beta), int(pad_id))
        return loss[0].clone()

    @staticmethod
    def backward(ctx, dloss):
        from . import lib as L
        lib = L.load()
        z, lab = ctx.saved_tensors
        N, V, Vp, alpha, beta, pad_id = ctx.meta
        dz = torch.empty((N, Vp), dtype=torch.float32, device=z.device)
        up = dloss.reshape(1).to(torch.float32).contiguous()
        stream = torch.cuda.current_stream(z.device).cuda_stream
        L.check(lib.vct_sce(z.data_ptr(), Vp, lab.data_ptr(), 1, N, 1, V, alpha, beta, pad_id, None, None, None,
                            dz.data_ptr(), L.F32, Vp, up.data_ptr(), stream), "vct_sce")
        return dz[:, :V], None, None, None, None
