"""ctypes binding of ``libsamvit_b200.so`` (C ABI in ``include/samvit_b200.h``).

The library must have been built (``__graft_entry__.build()`` / ``build.py``).  There is no fallback: if the
shared object is missing, loading raises, and every op requires a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsamvit_b200.so")

MODE_BF16, MODE_FP32 = 0, 1
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2


class SvbConfig(C.Structure):
    _fields_ = [
        ("img_size", C.c_int32), ("patch_size", C.c_int32), ("in_chans", C.c_int32), ("embed_dim", C.c_int32),
        ("depth", C.c_int32), ("num_heads", C.c_int32), ("mlp_dim", C.c_int32), ("window_size", C.c_int32),
        ("num_global", C.c_int32), ("global_idx", C.c_int32 * 16), ("fpn_dims", C.c_int32 * 4),
        ("ln_eps", C.c_float), ("gn_eps", C.c_float),
    ]


# every symbol include/samvit_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _sz, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_double
SYMBOLS = {
    "svb_last_error": (C.c_char_p, []),
    "svb_version": (_i, []),
    "svb_encoder_create": (_i, [C.POINTER(SvbConfig), C.POINTER(_vp)]),
    "svb_encoder_destroy": (None, [_vp]),
    "svb_encoder_load_param": (_i, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "svb_encoder_missing_params": (_i, [_vp]),
    "svb_encoder_workspace_bytes": (_sz, [_vp, _i, _i]),
    "svb_encoder_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "svb_encoder_workspace_bytes_hw": (_sz, [_vp, _i, _i, _i, _i]),
    "svb_encoder_forward_hw": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "svb_encoder_forward_x": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "svb_resize_pos_embed": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "svb_resize_rel_pos": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "svb_encoder_forward_u8": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "svb_stage_images_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "svb_encoder_forward_host": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "svb_encoder_enable_taps": (_i, [_vp, _i]),
    "svb_encoder_read_tap": (_i, [_vp, _i, _vp, _i64, _vp]),
    "svb_linear": (_i, [_i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _i, _vp]),
    "svb_linear_fused": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _i, _f, _vp, _i, _vp, _i, _i, _vp]),
    "svb_fold_layernorm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "svb_attention_tc": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "svb_attention_tc_phases": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "svb_rel_pack_rows": (_i, [_i, _i]),
    "svb_attention_debug_buffer": (_i, [_vp]),
    "svb_pack_rel_table": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "svb_fill_pad_rows": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "svb_fill_pad_rows_hw": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "svb_attention_window_hw": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "svb_attention_global_hw_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "svb_attention_global_hw": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "svb_add_cast": (_i, [_vp, _vp, _vp, _i, _i64, _vp]),
    "svb_layernorm": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "svb_attention": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "svb_im2col": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "svb_groupnorm_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i64, _f, _i, _vp]),
    "svb_ms_deform_attn_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "svb_ms_deform_attn_fused_forward": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "svb_profile_start": (_i, []),
    "svb_profile_stop": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "svb_launch_count": (C.c_int64, []),
    "svb_nchw_to_rows": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i64, _vp]),
    "svb_nchw_to_seq": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "svb_rows_to_nchw": (_i, [_vp, _i64, _vp, _i, _i, _i, _vp]),
    "svb_groupnorm_rows": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "svb_upsample_add_rows": (_i, [_vp, _i64, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "svb_im2col3x3_rows": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "svb_conv3x3_rows": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "svb_add_cast_bcast": (_i, [_vp, _vp, _i64, _vp, _i, _i64, _vp]),
    "svb_linear_nt": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "svb_fpn_conv3x3_rows": (_i, [_vp, _vp, _vp, _i, _f, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "svb_pass_schedule_model": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _d, _d, _vp, _i]),
    "svb_encoder_pass_schedule": (_i, [_vp, _i, _i, _i, _i, _vp, _i]),
    "svb_mask_threshold_heads_clear": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "svb_layernorm_post": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _f, _vp]),
    "svb_cls_token_recompute": (_i, [_vp, _i, _i, _i, _vp]),
    "svb_resize_bicubic_aa": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "svb_mask_threshold_heads": (_i, [_vp, _vp, _i, _i, _i64, _vp]),
    "svb_masked_cross_attention_workspace": (_i64, [_i, _i, _i, _i]),
    "svb_masked_cross_attention": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "svb_l2_normalize_rows": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _vp]),
    "svb_mask_clear_full_rows": (_i, [_vp, _i64, _i, _vp]),
    "svb_groupnorm_apply_nchw": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp]),
}

# the separate probe library (include/samvit_b200_probe.h): hardware probes used by tests / tools only
PROBE_LIB_PATH = os.path.join(HERE, "libsamvit_probe.so")
PROBE_SYMBOLS = {
    "svb_probe_last_error": (C.c_char_p, []),
    "svb_probe_mma_rate": (_i, [_i, _i, _i, _vp, _vp]),
    "svb_probe_tmem_rate": (_i, [_i, _i, _i, _vp, _vp]),
    "svb_probe_mma": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i] + [C.c_uint] * 6 + [_vp]),
}

_lib: Optional[C.CDLL] = None
_probe_lib: Optional[C.CDLL] = None


class SvbError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SvbError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built (run __graft_entry__.build()). "
                "There is no CPU / PyTorch fallback for this path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)       # raises AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def probe_lib() -> C.CDLL:
    global _probe_lib
    if _probe_lib is None:
        if not os.path.exists(PROBE_LIB_PATH):
            raise SvbError(f"{PROBE_LIB_PATH} is missing: run __graft_entry__.build()")
        l = C.CDLL(PROBE_LIB_PATH)
        for name, (res, args) in PROBE_SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _probe_lib = l
    return _probe_lib


def check_probe(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = probe_lib().svb_probe_last_error()
        raise SvbError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().svb_last_error()
        raise SvbError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
