"""Multi-scale deformable attention forward (scope row N1), host-side mirror of the reference operator.

``ms_deform_attn_forward`` has the argument list of ``MSDeformAttnFunction.forward``
(``/root/reference/modeling/vision/encoder/ops/functions/ms_deform_attn_func.py:34-40``) and calls the sm_100a kernel through
the C ABI (``svb_ms_deform_attn_forward``).  Forward only; CUDA tensors only; no fallback.
"""
from __future__ import annotations

import copy
import ctypes
import ctypes as C
import math

import torch
from torch import nn

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

    def _mode(self):
        if self.precision == "bf16":
            return cabi.MODE_BF16, torch.bfloat16
        if self.precision == "fp32":
            return cabi.MODE_FP32, torch.float32
        raise ValueError("precision must be 'bf16' or 'fp32'")

    def _core(self, q_in, x_in, reference_points, shapes, starts, input_padding_mask, N, Lq, S, resid=None):
        """q_in (N*Lq,C) / x_in (N*S,C) already in the GEMM operand dtype -> fp32 (N*Lq,C).  ms_deform_attn.py:97-124.
        With ``resid`` (fp32 (N*Lq,C)) the output projection accumulates IN PLACE, ``resid += output_proj(...)`` (the caller's
        ``src + dropout(src2)``, transformer_encoder_deform.py:126, in the GEMM's reduce-add epilogue), and ``resid`` is returned."""
        mode, adt = self._mode()
        bf16 = adt == torch.bfloat16
        dev = q_in.device
        C = self.d_model
        M, L, P = self.n_heads, self.n_levels, self.n_points
        D = C // M
        (wv, bv), (wq, bq), (wo, bo) = self._weights(dev, adt)
        value = self._linear(mode, x_in, wv, bv, torch.empty(N * S, C, dtype=adt, device=dev))                          # :97
        if input_padding_mask is not None:
            value.view(N, S, C).masked_fill_(input_padding_mask[..., None], 0.0)                                         # :98-99
        raw = self._linear(mode, q_in, wq, bq, torch.empty(N * Lq, 3 * M * L * P, dtype=torch.float32, device=dev))       # :101-102
        ref = reference_points.detach().to(torch.float32).contiguous()
        sampled = torch.empty(N * Lq, C, dtype=adt, device=dev)
        cabi.check(cabi.lib().svb_ms_deform_attn_fused_forward(
            value.data_ptr(), (ctypes.c_int32 * (2 * L))(*shapes), (ctypes.c_int32 * L)(*starts), ref.data_ptr(), int(ref.shape[-1]),
            raw.data_ptr(), sampled.data_ptr(), cabi.DTYPE_BF16 if bf16 else cabi.DTYPE_F32, N, S, M, D, L, Lq, P, cabi.stream_ptr()),
            "svb_ms_deform_attn_fused_forward")                                                                          # :103-122
        if resid is not None:
            cabi.check(cabi.lib().svb_linear(mode, sampled.data_ptr(), sampled.stride(0), wo.data_ptr(), wo.stride(0), N * Lq, C, C, bo.data_ptr(), 0,
                                             resid.data_ptr(), C, 0, resid.data_ptr(), cabi.DTYPE_F32, C, None, 0, 0, 0, cabi.stream_ptr()),
                       "svb_linear")                                                                                    # :124 (+ :126)
            return resid
        return self._linear(mode, sampled, wo, bo, torch.empty(N * Lq, C, dtype=torch.float32, device=dev))             # :124

    @staticmethod
    def _linear(mode, a, w, b, out, act=0):
        m, k = a.shape
        n = w.shape[0]
        cabi.check(cabi.lib().svb_linear(
            mode, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), m, n, k, b.data_ptr(), act, None, 0, 0, out.data_ptr(),
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
        if len(shapes) != 2 * L or len(starts) != L or sum(shapes[2 * i] * shapes[2 * i + 1] for i in range(L)) != S:
            raise AssertionError("the levels do not cover input_flatten")            # ms_deform_attn.py:95
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        mode, adt = self._mode()
        dev = query.device
        with torch.cuda.device(dev):
            x_in = input_flatten.detach().reshape(N * S, C).to(adt).contiguous()
            q_in = query.detach().reshape(N * Lq, C).to(adt).contiguous()
            out = self._core(q_in, x_in, reference_points, shapes, starts, input_padding_mask, N, Lq, S)
        return out.view(N, Lq, C).to(query.dtype)



# ----------------------------------------------------------------------------------------------------------------------
# The encoder around the module (scope row N1, widened to its caller): transformer_encoder_deform.py:23-161

def _layernorm(x, add, ln, out):
    """x (rows, dim) fp32, updated in place to x + add when ``add`` is given; out = LayerNorm(x) (fp32)."""
    rows, dim = x.shape
    cabi.check(cabi.lib().svb_layernorm(x.data_ptr(), add.data_ptr() if add is not None else None, ln._w32.data_ptr(), ln._b32.data_ptr(),
                                        out.data_ptr(), cabi.DTYPE_F32, rows, dim, float(ln.eps), cabi.stream_ptr()), "svb_layernorm")
    return out


_LN_POST_WIDTHS = (256, 512, 768, 1024, 1280)          # the widths svb_layernorm_post is built for


def _layernorm_post(x, ln, out, out_b=None, pos=None, out_q=None):
    """One pass: out = LayerNorm(x) (fp32), out_b = bf16(out), out_q = bf16(out + pos) (pos shared by the batch when it has fewer rows)."""
    rows, dim = x.shape
    cabi.check(cabi.lib().svb_layernorm_post(
        x.data_ptr(), None, ln._w32.data_ptr(), ln._b32.data_ptr(), out.data_ptr(), out_b.data_ptr() if out_b is not None else None,
        pos.data_ptr() if pos is not None else None, pos.numel() // dim if pos is not None else 0,
        out_q.data_ptr() if out_q is not None else None, rows, dim, float(ln.eps), cabi.stream_ptr()), "svb_layernorm_post")
    return out


def _add_cast(a, b, dtype):
    """cast(a [+ b]) of fp32 (rows, dim) streams to the GEMM operand dtype: ``with_pos_embed`` (:112-114) fused with the cast."""
    out = torch.empty(a.shape, dtype=dtype, device=a.device)
    odt = cabi.DTYPE_BF16 if dtype == torch.bfloat16 else cabi.DTYPE_F32
    if b is not None and b.numel() != a.numel():          # one embedding for every sample of the batch
        cabi.check(cabi.lib().svb_add_cast_bcast(a.data_ptr(), b.data_ptr(), b.numel(), out.data_ptr(), odt, a.numel(), cabi.stream_ptr()),
                   "svb_add_cast_bcast")
    else:
        cabi.check(cabi.lib().svb_add_cast(a.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(), odt, a.numel(),
                                           cabi.stream_ptr()), "svb_add_cast")
    return out


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    """Drop-in for ``MSDeformAttnTransformerEncoderLayer`` (``transformer_encoder_deform.py:91-136``): same constructor, sub-module
    names (``self_attn``, ``norm1``, ``linear1``, ``linear2``, ``norm2``; the dropouts are identities in the forward-only path) and
    ``forward`` arguments.  The residual stream stays fp32; ``src + pos`` is fused with the cast to the GEMM operand type, the residual
    add of the attention output with ``norm1``, the ReLU with ``linear1``'s epilogue and the FFN residual with ``linear2``'s."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("the B200 deformable encoder layer implements activation='relu' (the reference's configuration)")
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self._packed_sig = None

    @property
    def precision(self):
        return self.self_attn.precision

    @precision.setter
    def precision(self, p):
        self.self_attn.precision = p

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def _prepare(self, device, wdtype):
        ps = [self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias, self.norm1.weight, self.norm1.bias,
              self.norm2.weight, self.norm2.bias]
        sig = (str(device), wdtype) + tuple((p.data_ptr(), p._version) for p in ps)
        if sig != self._packed_sig:
            f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()      # noqa: E731
            self._w1, self._b1 = f(self.linear1.weight).to(wdtype), f(self.linear1.bias)
            self._w2, self._b2 = f(self.linear2.weight).to(wdtype), f(self.linear2.bias)
            for ln in (self.norm1, self.norm2):
                ln._w32, ln._b32 = f(ln.weight), f(ln.bias)
            self._packed_sig = sig

    def _forward_rows(self, x, pos, reference_points, shapes, starts, padding_mask, N, S, carry=None, last=True):
        """x (N*S, C) fp32, OVERWRITTEN (src + attention output, then the layer's output); pos fp32, (N*S, C) or (S, C) shared by the
        batch, or None -> the new (N*S, C) fp32 rows (in x's storage).  ``carry``: a dict the encoder loops hand from layer to layer —
        in the bf16 path the norm2 pass of a layer that is not the ``last`` also writes the next layer's GEMM operands (bf16(src) and
        bf16(src + pos)), so the two cast passes at the top of the next layer disappear."""
        attn = self.self_attn
        mode, adt = attn._mode()
        bf16 = adt == torch.bfloat16
        dev = x.device
        self._prepare(dev, adt)
        C = x.shape[1]
        if carry is not None and carry.get("x") is x:
            q_in, x_in = carry["q_in"], carry["x_in"]
        else:
            q_in = _add_cast(x, pos, adt) if (pos is not None or bf16) else x                                   # :125 with_pos_embed
            x_in = q_in if pos is None else (_add_cast(x, None, adt) if bf16 else x)
        attn._core(q_in, x_in, reference_points, shapes, starts, padding_mask, N, S, S, resid=x)                # :125-126 x = src + src2
        fused = bf16 and C in _LN_POST_WIDTHS                  # (other widths: LayerNorm + separate cast passes)
        if fused:                                                                                               # :127 norm1 (+ linear1's operand)
            yb = torch.empty(x.shape, dtype=adt, device=dev)
            y = _layernorm_post(x, self.norm1, torch.empty_like(x), yb)
        else:
            y = _layernorm(x, None, self.norm1, torch.empty_like(x))
            yb = _add_cast(y, None, adt) if bf16 else y
        hid = attn._linear(mode, yb, self._w1, self._b1, torch.empty(N * S, self._w1.shape[0], dtype=adt, device=dev), act=2)     # :117 relu(linear1)
        # src + linear2(..) accumulated IN PLACE into y (the GEMM's TMA reduce-add epilogue), then norm2 into the buffer x leaves behind
        cabi.check(cabi.lib().svb_linear(mode, hid.data_ptr(), hid.stride(0), self._w2.data_ptr(), self._w2.stride(0), N * S, C, hid.shape[1],
                                         self._b2.data_ptr(), 0, y.data_ptr(), C, 0, y.data_ptr(), cabi.DTYPE_F32, C, None, 0, 0, 0,
                                         cabi.stream_ptr()), "svb_linear")                                      # :117-118 src + linear2(..)
        if fused and carry is not None and not last:                                                             # :119 norm2 (+ the next layer's operands)
            nx = torch.empty(x.shape, dtype=adt, device=dev)
            nq = torch.empty(x.shape, dtype=adt, device=dev) if pos is not None else None
            _layernorm_post(y, self.norm2, x, nx, pos, nq)
            carry.update(x=x, x_in=nx, q_in=nq if nq is not None else nx)
            return x
        return _layernorm(y, None, self.norm2, x)                                                               # :119

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        if not src.is_cuda:
            raise RuntimeError("MSDeformAttnTransformerEncoderLayer (B200) has no CPU path: the inputs must be CUDA tensors")
        if torch.is_grad_enabled() and (src.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("MSDeformAttnTransformerEncoderLayer (B200) implements the forward pass only: call it under torch.no_grad()")
        N, S, C = src.shape
        shapes = [int(v) for v in spatial_shapes.reshape(-1).tolist()]
        starts = [int(v) for v in level_start_index.reshape(-1).tolist()]
        with torch.cuda.device(src.device):
            x = src.detach().reshape(N * S, C).to(torch.float32).clone()
            p = None if pos is None else pos.detach().reshape(N * S, C).to(torch.float32).contiguous()
            out = self._forward_rows(x, p, reference_points, shapes, starts, padding_mask, N, S)
        return out.view(N, S, C).to(src.dtype)


class MSDeformAttnTransformerEncoder(nn.Module):
    """``transformer_encoder_deform.py:133-161``: ``num_layers`` clones of the layer over one set of reference points."""

    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        # :141-153 — pixel centres of every level, normalised by the valid extent, per (image, query, level)
        pts = []
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            H_, W_ = int(H_), int(W_)
            ys = torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device)
            xs = torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device)
            ry = ys[:, None].expand(H_, W_).reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            rx = xs[None, :].expand(H_, W_).reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            pts.append(torch.stack((rx, ry), -1))
        return torch.cat(pts, 1)[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        if not src.is_cuda:
            raise RuntimeError("MSDeformAttnTransformerEncoder (B200) has no CPU path: the inputs must be CUDA tensors")
        if torch.is_grad_enabled() and (src.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("MSDeformAttnTransformerEncoder (B200) implements the forward pass only: call it under torch.no_grad()")
        N, S, C = src.shape
        shapes_hw = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
        shapes = [v for hw in shapes_hw for v in hw]
        starts = [int(v) for v in level_start_index.reshape(-1).tolist()]
        with torch.cuda.device(src.device):
            ref = self.get_reference_points(shapes_hw, valid_ratios.to(torch.float32), src.device)
            x = src.detach().reshape(N * S, C).to(torch.float32).clone()
            p = None if pos is None else pos.detach().reshape(N * S, C).to(torch.float32).contiguous()
            carry = {}
            for i, layer in enumerate(self.layers):
                x = layer._forward_rows(x, p, ref, shapes, starts, padding_mask, N, S, carry, i + 1 == len(self.layers))
        return x.view(N, S, C).to(src.dtype)


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    """``transformer_encoder_deform.py:23-88``: flattens the feature levels, adds the level embedding to the positional
    embeddings and runs the encoder.  Returns ``(memory, spatial_shapes, level_start_index)`` like the reference."""

    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1, activation="relu",
                 num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model, self.nhead = d_model, nhead
        layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels, nhead, enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._reset_parameters()

    def _reset_parameters(self):
        # :45-52
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)

    @property
    def precision(self):
        return self.encoder.layers[0].precision

    @precision.setter
    def precision(self, p):
        for layer in self.encoder.layers:
            layer.precision = p

    def forward(self, srcs, pos_embeds):
        # :63-88; the masks of the reference are all-False (:64), so valid_ratios == 1 and no value is masked
        src_flatten, pos_flatten, spatial_shapes = [], [], []
        for lvl, (src, pos_embed) in enumerate(zip(srcs, pos_embeds)):
            bs, c, h, w = src.shape
            spatial_shapes.append((h, w))
            src_flatten.append(src.detach().flatten(2).transpose(1, 2))
            pos_flatten.append(pos_embed.detach().flatten(2).transpose(1, 2) + self.level_embed[lvl].detach().view(1, 1, -1))
        src_flatten = torch.cat(src_flatten, 1)
        pos_flatten = torch.cat(pos_flatten, 1)
        dev = src_flatten.device
        spatial_shapes = torch.as_tensor(spatial_shapes, dtype=torch.long, device=dev)
        level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
        valid_ratios = torch.ones(src_flatten.shape[0], len(srcs), 2, dtype=torch.float32, device=dev)
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, pos_flatten, None)
        return memory, spatial_shapes, level_start_index
