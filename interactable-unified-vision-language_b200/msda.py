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


# ----------------------------------------------------------------------------------------------------------------------
import ctypes  # noqa: E402
import math  # noqa: E402

from torch import nn  # noqa: E402


class MSDeformAttn(nn.Module):
    """Drop-in for the reference's ``MSDeformAttn`` module (``ops/modules/ms_deform_attn.py:35-125``): same constructor, same
    parameters (``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj`` — so its ``state_dict`` loads), same
    ``forward`` arguments.  Forward only, CUDA only.  The four Linear layers run on the tcgen05 GEMM (``svb_linear``; the two query
    Linears as ONE GEMM over the concatenated weights), the softmax over (levels x points), the sampling-location arithmetic and
    the bilinear gather in one kernel (``svb_ms_deform_attn_fused_forward``): sampling_locations / attention_weights never exist in
    memory, and the value maps stay bf16 (the reference's kernel is fp32-only).

    ``precision``: "bf16" (tensor cores, fp32 accumulate) or "fp32" (validation mode: fp32 FMA GEMMs, fp32 values)."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        self.im2col_step = 128
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self.precision = "bf16"
        self._packed = None
        self._packed_sig = None
        self._reset_parameters()

    def _reset_parameters(self):
        # ops/modules/ms_deform_attn.py:67-80
        nn.init.constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.)
        nn.init.constant_(self.attention_weights.bias.data, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.)

    def _weights(self, device, wdtype):
        """GEMM operands: [value_proj, cat(sampling_offsets, attention_weights), output_proj], re-derived when a parameter changed."""
        ps = [self.value_proj.weight, self.value_proj.bias, self.sampling_offsets.weight, self.sampling_offsets.bias,
              self.attention_weights.weight, self.attention_weights.bias, self.output_proj.weight, self.output_proj.bias]
        sig = (str(device), wdtype) + tuple((p.data_ptr(), p._version) for p in ps)
        if sig != self._packed_sig:
            f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()      # noqa: E731
            wq = torch.cat([f(self.sampling_offsets.weight), f(self.attention_weights.weight)], 0)
            bq = torch.cat([f(self.sampling_offsets.bias), f(self.attention_weights.bias)], 0)
            self._packed = [(f(self.value_proj.weight).to(wdtype), f(self.value_proj.bias)), (wq.to(wdtype).contiguous(), bq.contiguous()),
                            (f(self.output_proj.weight).to(wdtype), f(self.output_proj.bias))]
            self._packed_sig = sig
        return self._packed

    @staticmethod
    def _linear(mode, a, w, b, out):
        m, k = a.shape
        n = w.shape[0]
        cabi.check(cabi.lib().svb_linear(
            mode, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), m, n, k, b.data_ptr(), 0, None, 0, 0, out.data_ptr(),
            cabi.DTYPE_BF16 if out.dtype == torch.bfloat16 else cabi.DTYPE_F32, out.stride(0), None, 0, 0, 0, cabi.stream_ptr()), "svb_linear")
        return out

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index, input_padding_mask=None):
        """query (N,Lq,C); reference_points (N,Lq,L,2|4) in [0,1]; input_flatten (N,S,C); input_spatial_shapes (L,2) = (H,W);
        input_level_start_index (L,); input_padding_mask (N,S) True = padding  ->  (N,Lq,C)   (ms_deform_attn.py:82-125)."""
        if not query.is_cuda:
            raise RuntimeError("MSDeformAttn (B200) has no CPU path: the inputs must be CUDA tensors")
        if torch.is_grad_enabled() and (query.requires_grad or input_flatten.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("MSDeformAttn (B200) implements the forward pass only: call it under torch.no_grad()")
        N, Lq, C = query.shape
        _, S, _ = input_flatten.shape
        M, L, P = self.n_heads, self.n_levels, self.n_points
        D = C // M
        shapes = [int(v) for v in input_spatial_shapes.reshape(-1).tolist()]
        starts = [int(v) for v in input_level_start_index.reshape(-1).tolist()]
        if sum(shapes[2 * i] * shapes[2 * i + 1] for i in range(L)) != S:
            raise AssertionError("the levels do not cover input_flatten")            # ms_deform_attn.py:95
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        bf16 = self.precision == "bf16"
        if not bf16 and self.precision != "fp32":
            raise ValueError("precision must be 'bf16' or 'fp32'")
        mode, adt = (cabi.MODE_BF16, torch.bfloat16) if bf16 else (cabi.MODE_FP32, torch.float32)
        dev = query.device
        with torch.cuda.device(dev):
            (wv, bv), (wq, bq), (wo, bo) = self._weights(dev, adt)
            x_in = input_flatten.detach().reshape(N * S, C).to(adt).contiguous()
            q_in = query.detach().reshape(N * Lq, C).to(adt).contiguous()
            value = self._linear(mode, x_in, wv, bv, torch.empty(N * S, C, dtype=adt, device=dev))                    # :97
            if input_padding_mask is not None:
                value.view(N, S, C).masked_fill_(input_padding_mask[..., None], 0.0)                                     # :98-99
            raw = self._linear(mode, q_in, wq, bq, torch.empty(N * Lq, 3 * M * L * P, dtype=torch.float32, device=dev))   # :101-102
            ref = reference_points.detach().to(torch.float32).contiguous()
            sampled = torch.empty(N * Lq, C, dtype=adt, device=dev)
            cabi.check(cabi.lib().svb_ms_deform_attn_fused_forward(
                value.data_ptr(), (ctypes.c_int32 * (2 * L))(*shapes), (ctypes.c_int32 * L)(*starts), ref.data_ptr(), int(ref.shape[-1]),
                raw.data_ptr(), sampled.data_ptr(), cabi.DTYPE_BF16 if bf16 else cabi.DTYPE_F32, N, S, M, D, L, Lq, P, cabi.stream_ptr()),
                "svb_ms_deform_attn_fused_forward")                                                                      # :103-122
            out = self._linear(mode, sampled, wo, bo, torch.empty(N * Lq, C, dtype=torch.float32, device=dev))           # :124
        return out.view(N, Lq, C).to(query.dtype)

