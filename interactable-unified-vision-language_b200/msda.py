"""Multi-scale deformable attention forward (scope row N1), host-side mirror of the reference operator.

``ms_deform_attn_forward`` has the argument list of ``MSDeformAttnFunction.forward``
(``/root/reference/modeling/vision/encoder/ops/functions/ms_deform_attn_func.py:34-40``) and calls the sm_100a kernel through
the C ABI (``svb_ms_deform_attn_forward``).  Forward only; CUDA tensors only; no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import cabi


def ms_deform_attn_forward(value: torch.Tensor, value_spatial_shapes: torch.Tensor, value_level_start_index: torch.Tensor,
                           sampling_locations: torch.Tensor, attention_weights: torch.Tensor, im2col_step: int = 64) -> torch.Tensor:
    """value (N,S,M,D) fp32 or bf16; value_spatial_shapes (L,2) = (H,W); value_level_start_index (L,);
    sampling_locations (N,Lq,M,L,P,2) in [0,1]; attention_weights (N,Lq,M,L,P)  ->  (N,Lq,M*D) in value's dtype.
    ``im2col_step`` is accepted for signature compatibility (it only batches the reference's launches)."""
    if not value.is_cuda:
        raise RuntimeError("ms_deform_attn_forward has no CPU path: value must be a CUDA tensor")
    if value.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("value must be float32 or bfloat16")
    if torch.is_grad_enabled() and any(t.requires_grad for t in (value, sampling_locations, attention_weights)):
        raise RuntimeError("ms_deform_attn_forward implements the forward pass only: call it under torch.no_grad()")
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = sampling_locations.shape
    if M2 != M or two != 2 or tuple(attention_weights.shape) != (N, Lq, M, L, P):
        raise ValueError("inconsistent shapes of value / sampling_locations / attention_weights")
    shapes = [int(v) for v in value_spatial_shapes.reshape(-1).tolist()]
    starts = [int(v) for v in value_level_start_index.reshape(-1).tolist()]
    if len(shapes) != 2 * L or len(starts) != L:
        raise ValueError("value_spatial_shapes must be (L,2) and value_level_start_index (L,)")
    v = value.contiguous()
    loc = sampling_locations.to(torch.float32).contiguous()
    w = attention_weights.to(torch.float32).contiguous()
    out = torch.empty(N, Lq, M * D, dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        cabi.check(cabi.lib().svb_ms_deform_attn_forward(
            v.data_ptr(), (C.c_int32 * (2 * L))(*shapes), (C.c_int32 * L)(*starts), loc.data_ptr(), w.data_ptr(), out.data_ptr(),
            cabi.DTYPE_BF16 if value.dtype == torch.bfloat16 else cabi.DTYPE_F32, N, S, M, D, L, Lq, P, cabi.stream_ptr()),
            "svb_ms_deform_attn_forward")
    return out
