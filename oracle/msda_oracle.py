"""CPU ORACLE for the multi-scale deformable attention forward (scope row N1).  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` and tools that check the CUDA kernel may import this file; the product never does.

Two restatements of the same operator:
  * ``ms_deform_attn_core`` follows the reference's own debug/test implementation ``ms_deform_attn_core_pytorch``
    (``/root/reference/modeling/vision/encoder/ops/functions/ms_deform_attn_func.py:52-72``): per level an
    ``F.grid_sample(bilinear, zeros, align_corners=False)`` of the head-major value map, weighted sum over levels x points;
  * ``ms_deform_attn_loops`` follows the reference's CUDA kernel ``ms_deformable_im2col_gpu_kernel`` and its bilinear tap
    (``ops/src/cuda/ms_deform_im2col_cuda.cuh:242-303`` and ``:18-69``) with explicit Python loops (small cases only).

Parity pin: ``tests/golden/make_golden_msda.py`` executes the SOURCE TEXT of the reference's ``ms_deform_attn_core_pytorch``
(the module itself cannot be imported: it requires the compiled ``MultiScaleDeformableAttention`` extension at import time,
``ms_deform_attn_func.py:21-29``) on seeded inputs — the toy geometry of the reference's ``ops/test.py:24-29`` and larger
ones with out-of-range sampling locations — and commits inputs + outputs as ``tests/golden/msda_*.npz``;
``tests/test_oracle.py`` checks both restatements against them.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def ms_deform_attn_core(value, spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D); spatial_shapes [(H,W)]*L; sampling_locations (N,Lq,M,L,P,2) (x,y) in [0,1];
    attention_weights (N,Lq,M,L,P)  ->  (N,Lq,M*D).   ms_deform_attn_func.py:52-72."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    levels = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * sampling_locations - 1                                            # :58
    sampled = []
    for lid, (h, w) in enumerate(shapes):
        v = levels[lid].flatten(2).transpose(1, 2).reshape(N * M, D, h, w)        # :62
        g = grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)                     # :64  (N*M, Lq, P, 2)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))   # :66-67
    aw = attention_weights.transpose(1, 2).reshape(N * M, 1, Lq, L * P)           # :70
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()                                       # :72


def ms_deform_attn_loops(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """Explicit-loop form of the CUDA kernel (ms_deform_im2col_cuda.cuh:242-303, bilinear :18-69).  fp64 accumulation."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    v = value.double()
    out = torch.zeros(N, Lq, M, D, dtype=torch.float64)
    for n in range(N):
        for q in range(Lq):
            for m in range(M):
                col = torch.zeros(D, dtype=torch.float64)
                for l in range(L):
                    H, W = int(spatial_shapes[l][0]), int(spatial_shapes[l][1])
                    start = int(level_start_index[l])
                    for p in range(P):
                        loc_w = float(sampling_locations[n, q, m, l, p, 0])
                        loc_h = float(sampling_locations[n, q, m, l, p, 1])
                        wt = float(attention_weights[n, q, m, l, p])
                        h_im, w_im = loc_h * H - 0.5, loc_w * W - 0.5                       # :286-287
                        if not (h_im > -1 and w_im > -1 and h_im < H and w_im < W):          # :289
                            continue
                        h_low, w_low = math.floor(h_im), math.floor(w_im)                    # :25-26
                        h_high, w_high = h_low + 1, w_low + 1
                        lh, lw = h_im - h_low, w_im - w_low
                        hh, hw = 1 - lh, 1 - lw

                        def tap(y, x):
                            return v[n, start + y * W + x, m]
                        val = torch.zeros(D, dtype=torch.float64)
                        if h_low >= 0 and w_low >= 0:                                        # :40
                            val += hh * hw * tap(h_low, w_low)
                        if h_low >= 0 and w_high <= W - 1:                                   # :46
                            val += hh * lw * tap(h_low, w_high)
                        if h_high <= H - 1 and w_low >= 0:                                   # :52
                            val += lh * hw * tap(h_high, w_low)
                        if h_high <= H - 1 and w_high <= W - 1:                              # :58
                            val += lh * lw * tap(h_high, w_high)
                        col += val * wt                                                      # :291
                out[n, q, m] = col
    return out.reshape(N, Lq, M * D)


def ms_deform_attn_module(sd, query, reference_points, input_flatten, spatial_shapes, padding_mask, n_heads, n_levels, n_points):
    """MSDeformAttn.forward (ops/modules/ms_deform_attn.py:82-125) on a plain state_dict ``sd`` with the reference's keys
    (sampling_offsets / attention_weights / value_proj / output_proj .weight / .bias)."""
    N, Lq, C = query.shape
    _, S, _ = input_flatten.shape
    M, L, P = n_heads, n_levels, n_points
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    value = input_flatten @ sd["value_proj.weight"].t() + sd["value_proj.bias"]                        # :97
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)                                         # :98-99
    value = value.view(N, S, M, C // M)
    off = (query @ sd["sampling_offsets.weight"].t() + sd["sampling_offsets.bias"]).view(N, Lq, M, L, P, 2)          # :101
    aw = (query @ sd["attention_weights.weight"].t() + sd["attention_weights.bias"]).view(N, Lq, M, L * P)           # :102
    aw = F.softmax(aw, -1).view(N, Lq, M, L, P)                                                                       # :103
    if reference_points.shape[-1] == 2:                                                                               # :105-108
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:                                                                             # :109-111
        loc = reference_points[:, :, None, :, None, :2] + off / P * reference_points[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError("Last dim of reference_points must be 2 or 4")
    out = ms_deform_attn_core(value, shapes, loc, aw)                                                                 # :117-122
    return out @ sd["output_proj.weight"].t() + sd["output_proj.bias"]                                                # :124


def deform_encoder_only(sd, srcs, pos_embeds, n_heads, n_points, n_layers):
    """MSDeformAttnTransformerEncoderOnly.forward (transformer_encoder_deform.py:63-88) over the encoder (:155-161) and its layers
    (:121-131), eval mode (dropouts are identities), on a plain state_dict with the reference's keys.  Pinned by
    tests/golden/deform_encoder_*.npz (tests/golden/make_golden_deform_encoder.py: the unmodified reference classes)."""
    L = len(srcs)
    shapes = [tuple(s.shape[2:]) for s in srcs]
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)                                                  # :71,79
    pos = torch.cat([p.flatten(2).transpose(1, 2) + sd["level_embed"][i].view(1, 1, -1) for i, p in enumerate(pos_embeds)], 1)   # :73-75,81
    N, S, C = src.shape
    # all-False masks (:64) -> valid_ratios == 1 (:54-61,84); reference points (:141-153)
    ref = []
    for h, w in shapes:
        ys = (torch.arange(h, dtype=src.dtype) + 0.5) / h
        xs = (torch.arange(w, dtype=src.dtype) + 0.5) / w
        ref.append(torch.stack((xs[None, :].expand(h, w).reshape(-1), ys[:, None].expand(h, w).reshape(-1)), -1))
    ref = torch.cat(ref, 0)[None, :, None, :].expand(N, S, L, 2)

    def ln(x, w, b):
        mu = x.mean(-1, keepdim=True)
        var = ((x - mu) ** 2).mean(-1, keepdim=True)
        return (x - mu) / torch.sqrt(var + 1e-5) * w + b

    out = src
    for i in range(n_layers):
        pre = f"encoder.layers.{i}."
        attn_sd = {k[len(pre) + len("self_attn."):]: v for k, v in sd.items() if k.startswith(pre + "self_attn.")}
        src2 = ms_deform_attn_module(attn_sd, out + pos, ref, out, shapes, None, n_heads, L, n_points)              # :125
        out = ln(out + src2, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])                                       # :126-127
        hid = torch.relu(out @ sd[pre + "linear1.weight"].t() + sd[pre + "linear1.bias"])                            # :117
        out = ln(out + hid @ sd[pre + "linear2.weight"].t() + sd[pre + "linear2.bias"], sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])   # :117-119
    return out
