"""CPU ORACLE for the MSDeformAttn pixel decoder forward (scope row N1).  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` and tools that check the CUDA path may import this file; the product never does.

``pixel_decoder`` restates ``MSDeformAttnPixelDecoder.forward``
(``/root/reference/modeling/vision/encoder/transformer_encoder_deform.py:315-359``) on a plain state_dict with the reference's keys,
in whatever dtype the inputs have (fp64 in the tests): 1x1 convolutions as matrix products over the channel dimension, explicit
GroupNorm(32) arithmetic, ``PositionEmbeddingSine`` (``modeling/modules/position_encoding.py:29-53``, normalize=True, all-False mask)
written out, the encoder through ``msda_oracle.deform_encoder_only``.  The detectron2 ``Conv2d`` wrapper is conv -> norm -> activation
and ``get_norm("GN", C)`` is ``GroupNorm(32, C)`` (detectron2 v0.6, absent from the reference tree).

Parity pin: ``tests/golden/pixel_decoder_*.npz`` = outputs of the UNMODIFIED reference class run on CPU
(``tests/golden/make_golden_pixel_decoder.py``); ``tests/test_oracle.py`` checks this file against them.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import msda_oracle


def group_norm(x, weight, bias, groups=32, eps=1e-5):
    """nn.GroupNorm on (N, C, H, W)."""
    n, c, h, w = x.shape
    g = x.reshape(n, groups, -1)
    mu = g.mean(-1, keepdim=True)
    var = ((g - mu) ** 2).mean(-1, keepdim=True)
    return ((g - mu) / torch.sqrt(var + eps)).reshape(n, c, h, w) * weight.view(1, c, 1, 1) + bias.view(1, c, 1, 1)


def conv1x1(x, weight, bias=None):
    y = torch.einsum("nchw,oc->nohw", x, weight.reshape(weight.shape[0], -1))
    return y if bias is None else y + bias.view(1, -1, 1, 1)


def position_embedding_sine(x, num_pos_feats, temperature=10000, scale=2 * math.pi):
    """position_encoding.py:29-53 with normalize=True and mask=None."""
    n, _, h, w = x.shape
    y_embed = torch.arange(1, h + 1, dtype=x.dtype)[None, :, None].expand(n, h, w)
    x_embed = torch.arange(1, w + 1, dtype=x.dtype)[None, None, :].expand(n, h, w)
    eps = 1e-6
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=x.dtype)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def pixel_decoder(sd, features, n_heads, n_layers, n_points=4):
    """-> (mask_features, [three multi-scale maps, lowest resolution first])."""
    conv_dim = sd["mask_features.weight"].shape[1]
    srcs, pos = [], []
    for idx, f in enumerate(["res5", "res4", "res3"]):                                                       # :319-322
        x = features[f]
        y = conv1x1(x, sd[f"input_proj.{idx}.0.weight"], sd[f"input_proj.{idx}.0.bias"])
        srcs.append(group_norm(y, sd[f"input_proj.{idx}.1.weight"], sd[f"input_proj.{idx}.1.bias"]))
        pos.append(position_embedding_sine(x, conv_dim // 2))
    tsd = {k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}
    y = msda_oracle.deform_encoder_only(tsd, srcs, pos, n_heads, n_points, n_layers)                         # :325
    bs = y.shape[0]
    out, start = [], 0
    for s in srcs:                                                                                           # :328-335
        h, w = s.shape[2:]
        out.append(y[:, start:start + h * w].transpose(1, 2).reshape(bs, -1, h, w))
        start += h * w
    x = features["res2"]                                                                                     # :341-351 (one FPN level)
    cur = conv1x1(x, sd["adapter_1.weight"], sd.get("adapter_1.bias"))
    if "adapter_1.norm.weight" in sd:
        cur = group_norm(cur, sd["adapter_1.norm.weight"], sd["adapter_1.norm.bias"])
    z = cur + F.interpolate(out[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
    z = F.conv2d(z, sd["layer_1.weight"], sd.get("layer_1.bias"), padding=1)
    if "layer_1.norm.weight" in sd:
        z = group_norm(z, sd["layer_1.norm.weight"], sd["layer_1.norm.bias"])
    out.append(torch.relu(z))
    return conv1x1(out[-1], sd["mask_features.weight"], sd["mask_features.bias"]), out[:3]                  # :353-359
