"""Engine attachment shared by the drop-in modules: one ``CaptionEngine`` per model (or per stand-alone
sub-module), built lazily on first use because the reference constructs on CPU and then calls
``.to(device)`` (train.py:210)."""
import os

import torch


def module_dims(video_encoder, cap_decoder) -> dict:
    dims = {}
    if video_encoder is not None:
        lay = video_encoder.transformer_encoder.layers
        dims.update(Din=video_encoder.unify[0].in_features, d=video_encoder.unify[0].out_features,
                    H_enc=lay[0].self_attn.num_heads, F_enc=lay[0].linear1.out_features, L_enc=len(lay))
    else:
        dims.update(Din=8, H_enc=1, F_enc=8, L_enc=0)
    if cap_decoder is not None:
        lay = cap_decoder.decoder.layers
        dims.update(d=cap_decoder.generator.in_features, H_dec=lay[0].self_attn.num_heads,
                    F_dec=lay[0].linear1.out_features, L_dec=len(lay), V=cap_decoder.generator.out_features,
                    pad_id=cap_decoder.pad_id, alpha=cap_decoder.sce_loss_alpha, dropout=cap_decoder.dropout_p)
    else:
        dims.update(H_dec=1, F_dec=8, L_dec=0, V=8, pad_id=0, alpha=1.0, dropout=video_encoder.dropout_p)
    return dims


def build_engine(video_encoder, cap_decoder, device, precision=None, gemm_impl=None):
    from vct.engine import CaptionEngine
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("vct_b200: the caption hot path runs on CUDA (sm_100a) only -- there is no CPU fallback; "
                           "construct the model with device=torch.device('cuda') and move it there")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    precision = precision or os.environ.get("VCT_PRECISION", "bf16")
    return CaptionEngine(video_encoder, cap_decoder, dims=module_dims(video_encoder, cap_decoder), device=device,
                         precision=precision, gemm_impl=gemm_impl)


def param_device(module) -> torch.device:
    return next(module.parameters()).device
