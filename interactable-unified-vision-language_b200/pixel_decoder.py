"""MSDeformAttn pixel decoder (scope row N1): drop-in for ``MSDeformAttnPixelDecoder``
(``/root/reference/modeling/vision/encoder/transformer_encoder_deform.py:164-359``), forward only, CUDA only, no fallback.

Same constructor keywords, same parameter names (``input_proj.N.{0,1}``, ``transformer.*``, ``mask_features``, ``adapter_1`` /
``layer_1`` with their ``.norm`` — the layout detectron2's ``Conv2d`` wrapper gives them), same
``forward(features) -> (mask_features, multi_scale_features)``.  Everything between the NCHW inputs and the NCHW outputs runs on rows
(NHWC): the 1x1 convolutions are tcgen05 GEMMs, the 3x3 ``output_conv`` an im2col + GEMM, GroupNorm(32) / bilinear upsample-add /
layout changes are the streaming kernels of ``csrc/pixdec.cu``, the transformer is ``msda.MSDeformAttnTransformerEncoderOnly``'s
layers.  detectron2 / fvcore are not needed: ``Conv2d(norm=GN)`` = conv -> GroupNorm(32, C) -> activation, ``c2_xavier_fill`` =
kaiming_uniform(a=1) + zero bias (their published definitions).
"""
from __future__ import annotations

import ctypes
import math
import os

import torch
from torch import nn

from . import cabi
from .msda import MSDeformAttnTransformerEncoder, MSDeformAttnTransformerEncoderOnly


class _Conv2d(nn.Conv2d):
    """Parameter holder with detectron2's ``Conv2d`` attribute layout (``.norm`` sub-module, ``.activation``)."""

    def __init__(self, cin, cout, kernel_size, padding=0, bias=True, norm=None, activation=None):
        super().__init__(cin, cout, kernel_size=kernel_size, stride=1, padding=padding, bias=bias)
        self.norm = norm
        self.activation = activation


def _c2_xavier_fill(m):
    nn.init.kaiming_uniform_(m.weight, a=1)
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)


def _odt(t):
    return cabi.DTYPE_BF16 if t == torch.bfloat16 else cabi.DTYPE_F32


class MSDeformAttnPixelDecoder(nn.Module):
    def __init__(self, input_shape=None, *, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                 transformer_enc_layers=6, conv_dim=512, mask_dim=512, norm="GN", transformer_in_features=("res3", "res4", "res5"),
                 common_stride=4):
        super().__init__()
        if norm not in ("GN", "", None):
            raise NotImplementedError("the B200 pixel decoder implements norm='GN' (step1.yaml) or none")
        self.in_features = ["res2", "res3", "res4", "res5"]                                   # :199-201
        self.feature_strides = [4, 8, 16, 32]
        self.feature_channels = [128, 256, 512, 1024]
        self.transformer_in_features = self.in_features[1:]                                   # :204-206
        transformer_in_channels = self.feature_channels[1:]
        self.transformer_feature_strides = self.feature_strides[1:]
        self.transformer_num_feature_levels = len(self.transformer_in_features)
        self.input_proj = nn.ModuleList([nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim))
                                         for c in transformer_in_channels[::-1]])               # :209-216, res5 first
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)                                   # :226-228
            nn.init.constant_(proj[0].bias, 0)
        self.transformer = MSDeformAttnTransformerEncoderOnly(d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
                                                              dim_feedforward=transformer_dim_feedforward,
                                                              num_encoder_layers=transformer_enc_layers,
                                                              num_feature_levels=self.transformer_num_feature_levels)   # :230-237
        self.num_pos_feats, self.temperature, self.scale = conv_dim // 2, 10000, 2 * math.pi  # PositionEmbeddingSine(N_steps, normalize=True)
        self.mask_dim = mask_dim
        self.mask_features = _Conv2d(conv_dim, mask_dim, 1)                                   # :243-250
        _c2_xavier_fill(self.mask_features)
        self.maskformer_num_feature_levels = 3
        self.common_stride = common_stride
        stride = min(self.transformer_feature_strides)
        self.num_fpn_levels = int(math.log2(stride) - math.log2(self.common_stride))          # :256-257
        use_bias = norm in ("", None)
        self._laterals, self._outputs = [], []
        for idx, cin in enumerate(self.feature_channels[:self.num_fpn_levels]):               # :263-288
            ln = nn.GroupNorm(32, conv_dim) if not use_bias else None
            on = nn.GroupNorm(32, conv_dim) if not use_bias else None
            lateral = _Conv2d(cin, conv_dim, 1, bias=use_bias, norm=ln)
            output = _Conv2d(conv_dim, conv_dim, 3, padding=1, bias=use_bias, norm=on, activation="relu")
            _c2_xavier_fill(lateral)
            _c2_xavier_fill(output)
            self.add_module("adapter_{}".format(idx + 1), lateral)
            self.add_module("layer_{}".format(idx + 1), output)
            self._laterals.append(lateral)
            self._outputs.append(output)
        self._laterals, self._outputs = self._laterals[::-1], self._outputs[::-1]              # :291-292, top-down order
        self.conv_dim = conv_dim
        self.implicit_conv = True          # bf16 path: 3x3 output_conv as an implicit GEMM (False: im2col operand + GEMM, for A/B runs)
        self.fuse_fpn = os.environ.get("SVB_FPN_FUSE", "1") != "0"   # bf16 path: the passes around output_conv fused (False: one kernel per reference op, A/B)
        self._cache_sig, self._pos_cache = None, {}

    # ------------------------------------------------------------------------------------------------------------------
    @property
    def precision(self):
        return self.transformer.precision

    @precision.setter
    def precision(self, p):
        self.transformer.precision = p

    def _prepare(self, device, wdtype):
        ps = [p for p in self.parameters()]
        sig = (str(device), wdtype) + tuple((p.data_ptr(), p._version) for p in ps)
        if sig == self._cache_sig:
            return
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()      # noqa: E731
        self._proj = []
        for proj in self.input_proj:
            conv, gn = proj[0], proj[1]
            self._proj.append((f(conv.weight).reshape(conv.out_channels, -1).to(wdtype).contiguous(), f(conv.bias), f(gn.weight), f(gn.bias),
                               gn.num_groups, float(gn.eps)))
        self._lat, self._out = [], []
        for lat, out in zip(self._laterals, self._outputs):
            self._lat.append((f(lat.weight).reshape(lat.out_channels, -1).to(wdtype).contiguous(), None if lat.bias is None else f(lat.bias),
                              None if lat.norm is None else (f(lat.norm.weight), f(lat.norm.bias), lat.norm.num_groups, float(lat.norm.eps))))
            w3 = f(out.weight).permute(0, 2, 3, 1).reshape(out.out_channels, -1)              # (Cout, ky, kx, Cin): the im2col column order
            self._out.append((w3.to(wdtype).contiguous(), None if out.bias is None else f(out.bias),
                              None if out.norm is None else (f(out.norm.weight), f(out.norm.bias), out.norm.num_groups, float(out.norm.eps))))
        mf = self.mask_features
        self._mf = (f(mf.weight).reshape(mf.out_channels, -1).to(wdtype).contiguous(), f(mf.bias))
        self._pos_cache = {}
        self._cache_sig = sig

    def _pos_rows(self, shapes, device):
        """Sine position embedding of every level (modules/position_encoding.py:29-53 with normalize=True on an all-False mask: it
        depends on the level's shape only) + the level embedding (:73-75), as fp32 rows (S, C) shared by the batch.  Cached per
        shape set; a few small tensor ops on the device at the first call (parameter preparation, like the weight packing)."""
        key = tuple(shapes)
        if key not in self._pos_cache:
            rows = []
            npf = self.num_pos_feats
            dim_t = torch.arange(npf, dtype=torch.float32, device=device)
            dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
            for lvl, (h, w) in enumerate(shapes):
                y = torch.arange(1, h + 1, dtype=torch.float32, device=device) / (h + 1e-6) * self.scale
                x = torch.arange(1, w + 1, dtype=torch.float32, device=device) / (w + 1e-6) * self.scale
                py = y[:, None] / dim_t
                px = x[:, None] / dim_t
                py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)     # (h, npf)
                px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)     # (w, npf)
                pos = torch.cat((py[:, None, :].expand(h, w, npf), px[None, :, :].expand(h, w, npf)), dim=2).reshape(h * w, 2 * npf)
                rows.append(pos + self.transformer.level_embed[lvl].detach().to(device=device, dtype=torch.float32)[None, :])
            self._pos_cache[key] = torch.cat(rows, 0).contiguous()
        return self._pos_cache[key]

    @staticmethod
    def _linear(mode, a, w, bias, out, act=0):
        m, k = a.shape
        n = w.shape[0]
        cabi.check(cabi.lib().svb_linear(
            mode, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), m, n, k, bias.data_ptr() if bias is not None else None, act, None, 0, 0,
            out.data_ptr(), _odt(out.dtype), out.stride(0), None, 0, 0, 0, cabi.stream_ptr()), "svb_linear")
        return out

    @staticmethod
    def _groupnorm(x, x_stride, gn, out, out_stride, B, HW, C, relu, ws):
        gamma, beta, groups, eps = gn
        cabi.check(cabi.lib().svb_groupnorm_rows(x.data_ptr(), x_stride, gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), _odt(out.dtype),
                                                 out_stride, B, HW, C, groups, eps, 1 if relu else 0, ws.data_ptr(), cabi.stream_ptr()),
                   "svb_groupnorm_rows")

    def forward(self, features, rows_out=False):
        """features: dict res2..res5 of (B, C, H, W) CUDA tensors (fp32 or bf16) -> (mask_features (B, mask_dim, H2, W2) fp32,
        [three (B, conv_dim, H, W) fp32 maps, lowest resolution first])   (transformer_encoder_deform.py:315-359).
        rows_out=True (not in the reference): the mask features stay in the layout and type the next GEMM reads — returns
        (None, multi_scale_features, {"mask_rows": (B * H2 * W2, mask_dim) rows, "mask_shape": (B, mask_dim, H2, W2)}) for
        ``XDecoderMaskPath.forward(..., mask_rows=, mask_shape=)``, skipping the NCHW round trip of the largest map."""
        x5 = features[self.transformer_in_features[-1]]
        if not x5.is_cuda:
            raise RuntimeError("MSDeformAttnPixelDecoder (B200) has no CPU path: the features must be CUDA tensors")
        if torch.is_grad_enabled() and (any(t.requires_grad for t in features.values()) or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("MSDeformAttnPixelDecoder (B200) implements the forward pass only: call it under torch.no_grad()")
        layer0 = self.transformer.encoder.layers[0]
        mode, adt = layer0.self_attn._mode()
        dev = x5.device
        lib = cabi.lib()
        st = cabi.stream_ptr
        C = self.conv_dim
        with torch.cuda.device(dev):
            self._prepare(dev, adt)
            B = x5.shape[0]
            lvls = [features[f] for f in self.transformer_in_features[::-1]]                   # :319-322, res5 -> res3
            shapes = [(int(t.shape[2]), int(t.shape[3])) for t in lvls]
            starts = [0]
            for h, w in shapes[:-1]:
                starts.append(starts[-1] + h * w)
            S = sum(h * w for h, w in shapes)
            ws = torch.empty(B * 64 * 2, dtype=torch.float64, device=dev)
            src = torch.empty(B * S, C, dtype=torch.float32, device=dev)                       # src_flatten (:79) as rows
            for idx, x in enumerate(lvls):
                x = x.detach().contiguous()
                if x.dtype not in (torch.float32, torch.bfloat16):
                    x = x.float()
                cin, (h, w) = x.shape[1], shapes[idx]
                rows = torch.empty(B * h * w, cin, dtype=adt, device=dev)
                cabi.check(lib.svb_nchw_to_rows(x.data_ptr(), _odt(x.dtype), rows.data_ptr(), _odt(adt), B, cin, h * w, 0, st()), "svb_nchw_to_rows")
                wp, bp, gw, gb, groups, eps = self._proj[idx]
                conv = self._linear(mode, rows, wp, bp, torch.empty(B * h * w, C, dtype=torch.float32, device=dev))     # input_proj[idx][0]
                self._groupnorm(conv, 0, (gw, gb, groups, eps), src[starts[idx]:], S * C, B, h * w, C, False, ws)       # input_proj[idx][1]
            pos = self._pos_rows(shapes, dev)
            # ---- transformer encoder (:325): all-False masks -> valid ratios 1 ----
            ref = MSDeformAttnTransformerEncoder.get_reference_points(shapes, torch.ones(B, len(shapes), 2, dtype=torch.float32, device=dev), dev)
            flat_shapes = [v for hw in shapes for v in hw]
            y = src
            carry, layers = {}, self.transformer.encoder.layers
            for i, layer in enumerate(layers):
                y = layer._forward_rows(y, pos, ref, flat_shapes, starts, None, B, S, carry, i + 1 == len(layers))
            # ---- extra FPN level(s) on the high-resolution features (:341-351) ----
            cur_shape, cur, cur_stride = shapes[-1], y[starts[-1]:], S * C                     # out[-1]: the finest transformer level
            for idx, f in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
                x = features[f].detach().contiguous()
                if x.dtype not in (torch.float32, torch.bfloat16):
                    x = x.float()
                cin, h, w = x.shape[1], int(x.shape[2]), int(x.shape[3])
                rows = torch.empty(B * h * w, cin, dtype=adt, device=dev)
                cabi.check(lib.svb_nchw_to_rows(x.data_ptr(), _odt(x.dtype), rows.data_ptr(), _odt(adt), B, cin, h * w, 0, st()), "svb_nchw_to_rows")
                wl, bl, nl = self._lat[idx]
                fpn = self._linear(mode, rows, wl, bl, torch.empty(B * h * w, C, dtype=torch.float32, device=dev))      # lateral_conv
                del rows
                wo, bo, no = self._out[idx]
                conv = torch.empty(B * h * w, C, dtype=torch.float32, device=dev)
                implicit = adt == torch.bfloat16 and w % 128 == 0 and C % 64 == 0 and self.implicit_conv
                if implicit and self.fuse_fpn:
                    # GroupNorm apply + `cur_fpn + F.interpolate(out[-1], ...)` (:348) + the zero-padded bf16 operand of the implicit-GEMM
                    # output_conv in ONE pass over the level (the two fp32 maps in between never exist)
                    pad = torch.empty(B * (h + 2) * (w + 2), C, dtype=torch.bfloat16, device=dev)
                    gwl, gbl, ggl, gel = nl if nl is not None else (None, None, 0, 0.0)
                    cabi.check(lib.svb_fpn_conv3x3_rows(fpn.data_ptr(), gwl.data_ptr() if gwl is not None else None,
                                                        gbl.data_ptr() if gbl is not None else None, ggl, gel, cur.data_ptr(), cur_stride,
                                                        cur_shape[0], cur_shape[1], wo.data_ptr(), bo.data_ptr() if bo is not None else None,
                                                        conv.data_ptr(), pad.data_ptr(), ws.data_ptr(), B, h, w, C, C, 1 if (no is None) else 0, st()),
                               "svb_fpn_conv3x3_rows")
                    del pad
                else:
                    if nl is not None:
                        self._groupnorm(fpn, 0, nl, fpn, 0, B, h * w, C, False, ws)
                    cabi.check(lib.svb_upsample_add_rows(cur.data_ptr(), cur_stride, fpn.data_ptr(), B, cur_shape[0], cur_shape[1], h, w, C, st()),
                               "svb_upsample_add_rows")                                         # cur_fpn + F.interpolate(out[-1], ...) (:348)
                    if implicit:
                        # output_conv (3x3) as an implicit GEMM: the A tiles are row-shifted boxes of a zero-padded bf16 copy of the map
                        # (68 MB per image at 256^2 x 512) instead of a 604 MB im2col operand
                        pad = torch.empty(B * (h + 2) * (w + 2), C, dtype=torch.bfloat16, device=dev)
                        cabi.check(lib.svb_conv3x3_rows(fpn.data_ptr(), wo.data_ptr(), bo.data_ptr() if bo is not None else None, conv.data_ptr(),
                                                        pad.data_ptr(), B, h, w, C, C, 1 if (no is None) else 0, st()), "svb_conv3x3_rows")
                        del pad
                    else:
                        per = max(1, min(B, (1 << 31) // (h * w * 9 * C * (2 if adt == torch.bfloat16 else 4))))     # images per im2col pass (<= 2 GB)
                        col = torch.empty(per * h * w, 9 * C, dtype=adt, device=dev)
                        for b0 in range(0, B, per):
                            nb = min(per, B - b0)
                            cabi.check(lib.svb_im2col3x3_rows(fpn[b0 * h * w:].data_ptr(), col.data_ptr(), _odt(adt), nb, h, w, C, st()), "svb_im2col3x3_rows")
                            self._linear(mode, col[:nb * h * w], wo, bo, conv[b0 * h * w:(b0 + nb) * h * w],
                                         act=2 if (no is None) else 0)                          # output_conv (3x3)
                        del col
                del fpn
                last_lvl = idx + 1 == self.num_fpn_levels
                cur_b = None
                if no is not None:
                    if last_lvl and adt == torch.bfloat16 and self.fuse_fpn:
                        # the last level's map is read by the mask_features convolution only: norm + F.relu write its bf16 operand directly
                        cur_b = torch.empty(B * h * w, C, dtype=adt, device=dev)
                        self._groupnorm(conv, 0, no, cur_b, 0, B, h * w, C, True, ws)
                    else:
                        self._groupnorm(conv, 0, no, conv, 0, B, h * w, C, True, ws)           # norm + F.relu
                cur_shape, cur, cur_stride = (h, w), conv, h * w * C
            # ---- outputs (:353-359) ----
            h, w = cur_shape
            wm, bm = self._mf
            cur_a = cur if (cur_stride == h * w * C) else None
            if cur_a is None:        # no FPN level: the finest transformer level itself, gathered densely
                cur_a = torch.empty(B * h * w, C, dtype=torch.float32, device=dev)
                cur_a.view(B, h * w, C).copy_(y.view(B, S, C)[:, starts[-1]:starts[-1] + h * w])
            if self.num_fpn_levels > 0 and cur_b is not None:
                a_in = cur_b
            else:
                a_in = cur_a if adt == torch.float32 else torch.empty(B * h * w, C, dtype=adt, device=dev)
                if adt != torch.float32:
                    cabi.check(lib.svb_add_cast(cur_a.data_ptr(), None, a_in.data_ptr(), _odt(adt), cur_a.numel(), st()), "svb_add_cast")
            mrows = self._linear(mode, a_in, wm, bm, torch.empty(B * h * w, self.mask_dim, dtype=adt if rows_out else torch.float32, device=dev))
            mask = None
            if not rows_out:
                mask = torch.empty(B, self.mask_dim, h, w, dtype=torch.float32, device=dev)
                cabi.check(lib.svb_rows_to_nchw(mrows.data_ptr(), 0, mask.data_ptr(), B, self.mask_dim, h * w, st()), "svb_rows_to_nchw")
            multi = []
            for idx, (lh, lw) in enumerate(shapes[:self.maskformer_num_feature_levels]):
                o = torch.empty(B, C, lh, lw, dtype=torch.float32, device=dev)
                cabi.check(lib.svb_rows_to_nchw(y[starts[idx]:].data_ptr(), S * C, o.data_ptr(), B, C, lh * lw, st()), "svb_rows_to_nchw")
                multi.append(o)
        if rows_out:
            return mask, multi, {"mask_rows": mrows, "mask_shape": (B, self.mask_dim, h, w)}
        return mask, multi
