"""ctypes binding of libvct_b200.so (include/vct.h).

The library is the product: there is no Python / PyTorch fallback.  ``load()`` raises if the
shared object is missing or does not export every declared symbol, and every op raises
``VctError`` on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "libvct_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU_FWD, ACT_GELU_BWD, ACT_GELU_FWD_F, ACT_MUL_AUX = 0, 1, 2, 3, 4
GEMM_SIMT, GEMM_TCGEN05, GEMM_TCGEN05_X3, GEMM_TCGEN05_X6 = 0, 1, 2, 3

vp, ll, i32, u32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_uint, C.c_float


class VctError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32),
        ("A", vp), ("a_dtype", i32), ("lda", ll), ("a_trans", i32),
        ("B", vp), ("b_dtype", i32), ("ldb", ll), ("b_trans", i32),
        ("C", vp), ("c_dtype", i32), ("ldc", ll),
        ("C2", vp), ("c2_dtype", i32), ("ldc2", ll),
        ("bias", vp),
        ("row_table", vp), ("row_period", i32),
        ("addend", vp), ("ld_addend", ll),
        ("act", i32),
        ("aux", vp), ("aux_dtype", i32), ("ld_aux", ll),
        ("drop_p", f32), ("rng_state", vp), ("site", u32),
        ("impl", i32),
        ("splitk_ws", vp), ("splitk_ws_floats", ll),
        ("split_ws", vp), ("split_ws_bytes", ll),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("B", i32), ("H", i32), ("Lq", i32), ("Lk", i32), ("dh", i32),
        ("dtype", i32),
        ("q", vp), ("q_ld", ll), ("k", vp), ("k_ld", ll), ("v", vp), ("v_ld", ll), ("o", vp), ("o_ld", ll),
        ("key_pad", vp), ("causal", i32), ("scale", f32),
        ("drop_p", f32), ("rng_state", vp), ("site", u32),
        ("probs", vp),
        ("d_o", vp), ("do_ld", ll), ("dq", vp), ("dq_ld", ll), ("dk", vp), ("dk_ld", ll), ("dv", vp), ("dv_ld", ll),
        ("q_bs", ll), ("k_bs", ll), ("v_bs", ll), ("o_bs", ll), ("do_bs", ll), ("dq_bs", ll), ("dk_bs", ll), ("dv_bs", ll),
        ("dbias", vp), ("dbias_partials", vp), ("dbias_counters", vp),
        ("row_stats", vp),
    ]


class MhaArgs(C.Structure):
    _fields_ = [
        ("B", i32), ("L", i32), ("Lk", i32), ("d", i32), ("H", i32),
        ("dtype", i32),
        ("x", vp), ("mem", vp), ("w_in", vp), ("b_in", vp),
        ("qkv", vp), ("kv", vp), ("kv_ready", i32),
        ("o", vp),
        ("key_pad", vp),
        ("drop_p", f32), ("rng_state", vp), ("site", u32),
        ("probs", vp),
        ("gemm_impl", i32),
        ("split_ws", vp), ("split_ws_bytes", ll),
    ]


# name -> (restype, argtypes); mirrors include/vct.h one to one
SIGNATURES = {
    "vct_version": (i32, []),
    "vct_last_error": (C.c_char_p, []),
    "vct_device_info": (i32, [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "vct_step_tick": (i32, [vp, vp, vp]),
    "vct_gemm": (i32, [C.POINTER(GemmArgs), vp]),
    "vct_gemm_grouped": (i32, [C.POINTER(GemmArgs), i32, vp]),
    "vct_gemm_split_workspace_bytes": (ll, [i32, i32, i32, i32, i32, i32]),
    "vct_split_bf16": (i32, [vp, ll, i32, i32, i32, i32, i32, vp, ll, vp]),
    "vct_gemm_tune": (i32, [i32, i32, i32]),
    "vct_gemm_trace": (i32, [vp]),
    "vct_prep_frames": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "vct_attn_fwd": (i32, [C.POINTER(AttnArgs), vp]),
    "vct_attn_bwd": (i32, [C.POINTER(AttnArgs), vp]),
    "vct_attn_enc_self_fwd": (i32, [C.POINTER(MhaArgs), vp]),
    "vct_attn_dec_self_fwd": (i32, [C.POINTER(MhaArgs), vp]),
    "vct_attn_dec_cross_fwd": (i32, [C.POINTER(MhaArgs), vp]),
    "vct_ln_residual_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, f32, vp, u32, vp]),
    "vct_ln_bwd_workspace_floats": (ll, [i32, i32]),
    "vct_ln_residual_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, f32, vp, u32, vp]),
    "vct_ln_bwd_reduce": (i32, [vp, i32, i32, vp, vp, vp, vp]),
    "vct_zero_rows": (i32, [vp, vp, i32, i32, vp]),
    "vct_stage_inputs": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, ll, vp]),
    "vct_embed_fwd": (i32, [vp, ll, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp, u32, vp]),
    "vct_embed_bwd": (i32, [vp, ll, vp, vp, i32, i32, i32, i32, i32, f32, vp, u32, vp]),
    "vct_embed_bwd_rows": (i32, [vp, vp, i32, i32, i32, f32, vp, u32, vp]),
    "vct_embed_bwd_det": (i32, [vp, ll, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "vct_embed_sort": (i32, [vp, ll, i32, i32, i32, i32, vp, vp]),
    "vct_embed_segment_sum": (i32, [vp, vp, vp, i32, i32, vp]),
    "vct_embed_mark": (i32, [vp, ll, i32, i32, i32, i32, vp, vp, vp]),
    "vct_adam_rows": (i32, [vp, vp, vp, vp, vp, i32, i32, vp, f32, vp, vp, i32, vp]),
    "vct_embed_zero": (i32, [vp, ll, vp, i32, i32, i32, i32, vp]),
    "vct_sce": (i32, [vp, ll, vp, ll, i32, i32, i32, f32, f32, i32, vp, vp, vp, vp, i32, ll, vp, vp]),
    "vct_sce_typed": (i32, [vp, i32, ll, vp, ll, i32, i32, i32, f32, f32, i32, vp, vp, vp, vp, i32, ll, vp, vp]),
    "vct_colsum_workspace_floats": (ll, [i32, i32]),
    "vct_colsum": (i32, [vp, i32, ll, i32, i32, vp, vp, vp, vp]),
    "vct_adam": (i32, [vp, vp, i32, vp, vp, vp, ll, vp, f32, vp]),
    "vct_cast": (i32, [vp, vp, i32, ll, vp]),
    "vct_argmax_append": (i32, [vp, ll, i32, i32, vp, ll, i32, i32, vp, vp, vp]),
    "vct_dropout_mask": (i32, [vp, ll, f32, vp, u32, vp]),
    "vct_comm_create": (i32, [i32, i32, ll, i32, C.POINTER(vp)]),
    "vct_comm_base": (vp, [vp]),
    "vct_comm_ipc_handle": (i32, [vp, C.POINTER(C.c_ubyte)]),
    "vct_comm_connect": (i32, [vp, C.c_char_p]),
    "vct_comm_connect_in_process": (i32, [vp, C.POINTER(vp), C.POINTER(i32)]),
    "vct_peer_allreduce_bf16": (i32, [vp, ll, ll, ll, i32, vp]),
    "vct_peer_allgather": (i32, [vp, ll, ll, i32, vp]),
    "vct_comm_status": (i32, [vp]),
    "vct_comm_destroy": (i32, [vp]),
}

_lib = None


def load(path: str | None = None) -> C.CDLL:
    """Load the shared object and bind every symbol of include/vct.h (fails loudly otherwise)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("VCT_LIB", LIB_PATH)
    if not os.path.isfile(p):
        raise VctError(f"{p} not found: build it with `python video-captioning-transformer_b200/build.py` "
                       f"(there is no CPU / PyTorch fallback for this path)")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise VctError(f"{p} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vct_last_error()
        raise VctError(f"{what or 'vct'} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()
