"""Mask branch of the X-Decoder prediction heads (scope row N4, first slice), host-side mirror of
``XDecoder.forward_prediction_heads`` (``/root/reference/modeling/interface/xdecoder.py:429-494``) for the inference path with
``task_switch['mask']`` on: ``decoder_norm`` -> class-token recompute -> ``mask_embed`` MLP -> mask logits
``einsum("bqc,bchw->bqhw")`` -> antialiased bicubic resize to the attention-mask size -> ``sigmoid < 0.5`` per head.
Forward only, CUDA only, no fallback.  The class logits (``lang_encoder.compute_similarity``), boxes and captions belong to the text
side of the decoder and are not part of this slice.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import cabi


class _MLP(nn.Module):
    """Parameter holder with the key layout of the reference's ``MLP`` (``interface/modules.py:188-201``: ``layers.N.weight/bias``)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))


def _odt(t):
    return cabi.DTYPE_BF16 if t == torch.bfloat16 else cabi.DTYPE_F32


class MaskPredictionHead(nn.Module):
    """``decoder_norm`` and ``mask_embed`` carry the names they have inside ``XDecoder`` (``xdecoder.py:109,133``), so the matching
    entries of its ``state_dict`` load with ``strict=True``.  ``precision``: "bf16" (tcgen05 GEMMs) or "fp32" (validation mode)."""

    def __init__(self, hidden_dim=512, mask_dim=512, num_queries=101, nheads=8, dim_proj=None, bbox=False, caption=False):
        """``dim_proj``: width of the class / caption embedding (``class_embed``, xdecoder.py:135); with it the head also returns
        ``outputs_class`` (when text embeddings are passed to forward), ``outputs_caption`` (``caption=True``, the reference's
        task_switch['caption']) and, with ``bbox=True`` (task_switch['bbox']), ``outputs_bbox`` from ``bbox_embed`` (xdecoder.py:139)."""
        super().__init__()
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.mask_embed = _MLP(hidden_dim, hidden_dim, mask_dim, 3)
        self.num_queries, self.num_heads = num_queries, nheads
        self.caption = bool(caption)
        if dim_proj is not None:
            self.class_embed = nn.Parameter(torch.empty(hidden_dim, dim_proj))
            nn.init.trunc_normal_(self.class_embed, std=0.02)                              # xdecoder.py:136
        else:
            self.class_embed = None
        self.bbox_embed = _MLP(hidden_dim, hidden_dim, 4, 3) if bbox else None             # xdecoder.py:139
        self.precision = "bf16"
        self._sig = None

    def _prepare(self, device, wdtype):
        ps = list(self.parameters())
        sig = (str(device), wdtype) + tuple((p.data_ptr(), p._version) for p in ps)
        if sig != self._sig:
            f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()      # noqa: E731
            self._ln = (f(self.decoder_norm.weight), f(self.decoder_norm.bias))
            self._mlp = [(f(l.weight).to(wdtype).contiguous(), f(l.bias)) for l in self.mask_embed.layers]
            self._box = [(f(l.weight).to(wdtype).contiguous(), f(l.bias)) for l in self.bbox_embed.layers] if self.bbox_embed is not None else None
            # x @ class_embed as a GEMM against the [dim_proj, hidden] operand
            self._cls = f(self.class_embed).t().contiguous().to(wdtype).contiguous() if self.class_embed is not None else None
            self._sig = sig

    @staticmethod
    def _linear(mode, a, w, bias, out, act=0):
        m, k = a.shape
        cabi.check(cabi.lib().svb_linear(mode, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), m, w.shape[0], k,
                                         bias.data_ptr() if bias is not None else None, act, None, 0, 0, out.data_ptr(), _odt(out.dtype),
                                         out.stride(0), None, 0, 0, 0, cabi.stream_ptr()), "svb_linear")
        return out

    def mask_rows(self, mask_features):
        """mask_features (B, Cm, H, W) -> (B * H * W, Cm) rows in the GEMM operand type: the layout the mask-logit GEMM reads.  A caller that
        runs the head several times on the same features (one call per decoder layer) converts once and passes `rows=`."""
        adt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        mf = mask_features.detach().contiguous()
        if mf.dtype not in (torch.float32, torch.bfloat16):
            mf = mf.float()
        B, Cm, H, W = mf.shape
        rows = torch.empty(B * H * W, Cm, dtype=adt, device=mf.device)
        cabi.check(cabi.lib().svb_nchw_to_rows(mf.data_ptr(), _odt(mf.dtype), rows.data_ptr(), _odt(adt), B, Cm, H * W, 0, cabi.stream_ptr()),
                   "svb_nchw_to_rows")
        return rows

    def class_box_outputs(self, y, a, B, Q, mode, adt, text_embeddings=None, logit_scale=None):
        """The outputs of forward_prediction_heads beside the mask (xdecoder.py:452-484) from the decoder rows after the class-token
        recompute (``y`` fp32, ``a`` in the GEMM operand type): class embeddings (:453), class logits = the language encoder's
        compute_similarity against ``text_embeddings`` (K, dim_proj) with ``logit_scale`` (vlpencoder.py:239-245), boxes (:478),
        caption embeddings (:482)."""
        lib, st, dev = cabi.lib(), cabi.stream_ptr, y.device
        res = {"outputs_class": None, "outputs_bbox": None, "outputs_caption": None}
        if self._cls is not None:
            ce = self._linear(mode, a, self._cls, None, torch.empty(B * Q, self._cls.shape[0], dtype=torch.float32, device=dev))
            if self.caption:
                res["outputs_caption"] = ce.view(B, Q, -1)
            if text_embeddings is not None:
                K, DP = text_embeddings.shape
                scale = float(torch.as_tensor(logit_scale).exp()) if logit_scale is not None else 1.0
                v = torch.empty(B * Q, DP, dtype=adt, device=dev)
                cabi.check(lib.svb_l2_normalize_rows(ce.data_ptr(), v.data_ptr(), _odt(adt), B * Q, DP, 1e-7, scale, st()), "svb_l2_normalize_rows")
                t = text_embeddings.detach().to(device=dev, dtype=adt).contiguous()
                Kp = (K + 7) // 8 * 8                                                    # 16-byte rows for the GEMM's stores
                logits = torch.empty(B * Q, Kp, dtype=torch.float32, device=dev)
                self._linear(mode, v, t, None, logits[:, :K] if Kp == K else logits.as_strided((B * Q, K), (Kp, 1)))
                res["outputs_class"] = logits[:, :K].reshape(B, Q, K) if Kp != K else logits.view(B, Q, K)
        if self._box is not None:
            x = a
            n = len(self._box)
            for i, (w, b) in enumerate(self._box):
                last = i == n - 1
                odt = torch.float32 if last else adt
                Np = 8 if last else w.shape[0]                                            # the 4 box values live in an 8-wide row
                buf = torch.empty(B * Q, Np, dtype=odt, device=dev)
                x = self._linear(mode, x, w, b, buf.as_strided((B * Q, w.shape[0]), (Np, 1)), act=0 if last else 2)
            res["outputs_bbox"] = x.reshape(B, Q, 4)
        return res

    def forward(self, output, mask_features, attn_mask_target_size, rows=None, text_embeddings=None, logit_scale=None, clear_full_rows=False):
        """``clear_full_rows`` (not in the reference's signature): also apply the first statement of the NEXT decoder layer to the
        attention mask (xdecoder.py:267: rows whose every key is masked are cleared) while it is written.
        output (Q, B, C) — the decoder's query states as ``forward_prediction_heads`` receives them; mask_features (B, Cm, H, W) fp32 or
        bf16; attn_mask_target_size (h, w)  ->  {"outputs_mask": (B, Q, H, W) fp32, "attn_mask": (B * heads, Q, h * w) bool} and, when the
        head was built with ``dim_proj`` / ``bbox``, "outputs_class" / "outputs_bbox" / "outputs_caption" (see class_box_outputs)."""
        if not output.is_cuda:
            raise RuntimeError("MaskPredictionHead (B200) has no CPU path: the inputs must be CUDA tensors")
        if torch.is_grad_enabled() and (output.requires_grad or (torch.is_tensor(mask_features) and mask_features.requires_grad)
                                        or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("MaskPredictionHead (B200) implements the forward pass only: call it under torch.no_grad()")
        Q, B, C = output.shape
        if Q != self.num_queries:
            raise ValueError("this slice implements the segmentation path: `output` must hold exactly num_queries rows (no caption tokens)")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        mode, adt = (cabi.MODE_BF16, torch.bfloat16) if self.precision == "bf16" else (cabi.MODE_FP32, torch.float32)
        dev = output.device
        lib, st = cabi.lib(), cabi.stream_ptr
        Bm, Cm, H, W = tuple(mask_features.shape) if torch.is_tensor(mask_features) else tuple(mask_features)   # a shape when `rows` is given
        oh, ow = int(attn_mask_target_size[0]), int(attn_mask_target_size[1])
        with torch.cuda.device(dev):
            self._prepare(dev, adt)
            x = output.detach().to(torch.float32).transpose(0, 1).contiguous().view(B * Q, C)         # :431 (the transpose, a 200 KB copy)
            y = torch.empty_like(x)
            cabi.check(lib.svb_layernorm(x.data_ptr(), None, self._ln[0].data_ptr(), self._ln[1].data_ptr(), y.data_ptr(), cabi.DTYPE_F32,
                                         B * Q, C, float(self.decoder_norm.eps), st()), "svb_layernorm")               # :430 decoder_norm
            cabi.check(lib.svb_cls_token_recompute(y.data_ptr(), B, Q, C, st()), "svb_cls_token_recompute")            # :440-450
            a = y
            if adt != torch.float32:
                a = torch.empty(B * Q, C, dtype=adt, device=dev)
                cabi.check(lib.svb_add_cast(y.data_ptr(), None, a.data_ptr(), _odt(adt), y.numel(), st()), "svb_add_cast")
            extra = self.class_box_outputs(y, a, B, Q, mode, adt, text_embeddings, logit_scale)                      # :452-455, 476-482
            n = len(self._mlp)
            for i, (w, b) in enumerate(self._mlp):                                                                    # :458 mask_embed
                last = i == n - 1
                a = self._linear(mode, a, w, b, torch.empty(B * Q, w.shape[0], dtype=adt, device=dev), act=0 if last else 2)
            if rows is None or rows.dtype != adt:
                rows = self.mask_rows(mask_features)
            masks = torch.empty(B, Q, H, W, dtype=torch.float32, device=dev)
            for b in range(B):                                                                                        # :459 "bqc,bchw->bqhw"
                if mode == cabi.MODE_BF16 and (H * W) % 8 == 0:
                    # positions along the GEMM's M (no tile rows spent on padding 101 queries to 256), result stored query-major
                    rb, ab = rows[b * H * W:(b + 1) * H * W], a[b * Q:(b + 1) * Q]
                    cabi.check(lib.svb_linear_nt(rb.data_ptr(), rb.stride(0), ab.data_ptr(), ab.stride(0), H * W, Q, rb.shape[1], None,
                                                 masks[b].data_ptr(), H * W, st()), "svb_linear_nt")
                else:
                    self._linear(mode, a[b * Q:(b + 1) * Q], rows[b * H * W:(b + 1) * H * W], None, masks[b].view(Q, H * W))
            tmp = torch.empty(B * Q * H * ow, dtype=torch.float32, device=dev)
            small = torch.empty(B, Q * oh * ow, dtype=torch.float32, device=dev)
            cabi.check(lib.svb_resize_bicubic_aa(masks.data_ptr(), tmp.data_ptr(), small.data_ptr(), B * Q, H, W, oh, ow, st()),
                       "svb_resize_bicubic_aa")                                                                        # :463
            attn = torch.empty(B * self.num_heads, Q, oh * ow, dtype=torch.bool, device=dev)
            if clear_full_rows:
                cabi.check(lib.svb_mask_threshold_heads_clear(small.data_ptr(), attn.data_ptr(), B, self.num_heads, Q, oh * ow, st()),
                           "svb_mask_threshold_heads_clear")                                                           # :467-470 + :267
            else:
                cabi.check(lib.svb_mask_threshold_heads(small.data_ptr(), attn.data_ptr(), B, self.num_heads, Q * oh * ow, st()),
                           "svb_mask_threshold_heads")                                                                 # :467-470
        return {"outputs_mask": masks, "attn_mask": attn, "attn_logits": small.view(B, Q, oh, ow), **extra}


class CrossAttentionLayer(nn.Module):
    """Drop-in for the reference's ``CrossAttentionLayer`` (``interface/modules.py:72-131``, post-norm path ``forward_post``): same
    constructor and parameter names (``multihead_attn.in_proj_weight / in_proj_bias / out_proj.*``, ``norm.*`` — an
    ``nn.MultiheadAttention`` is kept as the parameter holder), same ``forward(tgt, memory, memory_mask, ..., pos, query_pos)``.
    Forward only, CUDA only.  The three input projections and the output projection run on the GEMM (``svb_linear``), ``tensor + pos``
    is fused with the cast to the operand type, the masked softmax(QK^T)V is ``svb_masked_cross_attention``, the residual add is fused
    into the LayerNorm.  Returns ``(tgt, None)``: the head-averaged attention map the reference also returns is not materialised."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        if normalize_before:
            raise NotImplementedError("the B200 CrossAttentionLayer implements the post-norm path (PRE_NORM: False in step1.yaml)")
        self._mha_name = "multihead_attn"
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.normalize_before = normalize_before
        self.nhead = nhead
        self.precision = "bf16"
        self._sig = None
        for p in self.parameters():                                                    # :87-90
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def _prepare(self, device, wdtype):
        ps = list(self.parameters())
        sig = (str(device), wdtype) + tuple((p.data_ptr(), p._version) for p in ps)
        if sig != self._sig:
            f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()      # noqa: E731
            mha = getattr(self, self._mha_name)
            w, b = f(mha.in_proj_weight), f(mha.in_proj_bias)
            c = w.shape[1]
            self._wq, self._wk, self._wv = (w[i * c:(i + 1) * c].to(wdtype).contiguous() for i in range(3))
            self._bq, self._bk, self._bv = (b[i * c:(i + 1) * c].contiguous() for i in range(3))
            self._wo, self._bo = f(mha.out_proj.weight).to(wdtype).contiguous(), f(mha.out_proj.bias)
            self.norm._w32, self.norm._b32 = f(self.norm.weight), f(self.norm.bias)
            self._sig = sig

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None, memory_operands=None):
        """tgt (Q, B, C), memory (HW, B, C), memory_mask (B * heads, Q, HW) bool (True = not allowed), pos / query_pos like memory / tgt.
        ``memory_operands`` (not in the reference's signature): a dict a caller may pass to every layer that reads the SAME memory / pos
        (the X-Decoder visits each feature level three times, xdecoder.py:262): the first call stores the GEMM operand copies of
        ``memory + pos`` and ``memory`` in it, the later ones reuse them instead of casting the level again."""
        if not tgt.is_cuda:
            raise RuntimeError("CrossAttentionLayer (B200) has no CPU path: the inputs must be CUDA tensors")
        if memory_key_padding_mask is not None:
            raise NotImplementedError("memory_key_padding_mask is not used by the X-Decoder (xdecoder.py:259-263) and not implemented")
        if torch.is_grad_enabled() and (tgt.requires_grad or memory.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("CrossAttentionLayer (B200) implements the forward pass only: call it under torch.no_grad()")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        mode, adt = (cabi.MODE_BF16, torch.bfloat16) if self.precision == "bf16" else (cabi.MODE_FP32, torch.float32)
        Q, B, C = tgt.shape
        HW = memory.shape[0]
        h = self.nhead
        dev = tgt.device
        lib, st = cabi.lib(), cabi.stream_ptr
        lin = MaskPredictionHead._linear

        def cast(a, b):
            out = torch.empty(a.shape, dtype=adt, device=dev)
            cabi.check(lib.svb_add_cast(a.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(), _odt(adt), a.numel(), st()), "svb_add_cast")
            return out

        with torch.cuda.device(dev):
            self._prepare(dev, adt)
            x = tgt.detach().to(torch.float32).contiguous().view(Q * B, C).clone()
            mem = memory.detach().to(torch.float32).contiguous().view(HW * B, C)
            qp = None if query_pos is None else query_pos.detach().to(torch.float32).contiguous().view(Q * B, C)
            kp = None if pos is None else pos.detach().to(torch.float32).contiguous().view(HW * B, C)
            q_in = cast(x, qp) if (qp is not None or adt != torch.float32) else x                                  # :100 with_pos_embed(tgt, query_pos)
            if memory_operands is not None and memory_operands.get("key") == (adt, HW * B, C):
                k_in, v_in = memory_operands["k_in"], memory_operands["v_in"]
            else:
                k_in = cast(mem, kp) if (kp is not None or adt != torch.float32) else mem                          # :101 with_pos_embed(memory, pos)
                v_in = k_in if kp is None else (cast(mem, None) if adt != torch.float32 else mem)                  # :102 value = memory
                if memory_operands is not None:
                    memory_operands.update(key=(adt, HW * B, C), k_in=k_in, v_in=v_in)
            q = lin(mode, q_in, self._wq, self._bq, torch.empty(Q * B, C, dtype=adt, device=dev))
            k = lin(mode, k_in, self._wk, self._bk, torch.empty(HW * B, C, dtype=adt, device=dev))
            v = lin(mode, v_in, self._wv, self._bv, torch.empty(HW * B, C, dtype=adt, device=dev))
            mask = None
            if memory_mask is not None:
                if memory_mask.dtype != torch.bool or tuple(memory_mask.shape) != (B * h, Q, HW):
                    raise ValueError("memory_mask must be a bool tensor of shape (batch * heads, queries, keys)")
                mask = memory_mask.contiguous()
            nws = int(lib.svb_masked_cross_attention_workspace(Q, HW, B, h))
            ws = torch.empty(nws, dtype=torch.float32, device=dev)
            att = torch.empty(Q * B, C, dtype=adt, device=dev)
            cabi.check(lib.svb_masked_cross_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), _odt(adt), mask.data_ptr() if mask is not None else None,
                                                      att.data_ptr(), ws.data_ptr(), nws, Q, HW, B, h, C // h, st()), "svb_masked_cross_attention")
            tgt2 = lin(mode, att, self._wo, self._bo, torch.empty(Q * B, C, dtype=torch.float32, device=dev))     # out_proj
            out = torch.empty(Q * B, C, dtype=torch.float32, device=dev)
            cabi.check(lib.svb_layernorm(x.data_ptr(), tgt2.data_ptr(), self.norm._w32.data_ptr(), self.norm._b32.data_ptr(), out.data_ptr(),
                                         cabi.DTYPE_F32, Q * B, C, float(self.norm.eps), st()), "svb_layernorm")          # :104-105
        return out.view(Q, B, C).to(tgt.dtype), None


class SelfAttentionLayer(CrossAttentionLayer):
    """Drop-in for the reference's ``SelfAttentionLayer`` (``interface/modules.py:14-69``, post-norm path): ``q = k = tgt + query_pos``,
    ``value = tgt``, ``attn_mask = tgt_mask`` — the same kernels as the cross-attention layer with the queries as the memory.  The
    attention parameters live under ``self_attn`` (the reference's own ``MultiheadAttention`` uses torch's parameter names)."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__(d_model, nhead, dropout, activation, normalize_before)
        self.self_attn = self.multihead_attn
        del self.multihead_attn
        self._mha_name = "self_attn"

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        if tgt_key_padding_mask is not None:
            raise NotImplementedError("tgt_key_padding_mask is not used by the X-Decoder (xdecoder.py:271-275) and not implemented")
        if tgt_mask is not None and tgt_mask.dtype != torch.bool:
            raise NotImplementedError("only boolean attention masks (True = not allowed) are implemented")
        out, _ = super().forward(tgt, tgt, memory_mask=tgt_mask, pos=query_pos, query_pos=query_pos)
        return out


class FFNLayer(nn.Module):
    """Drop-in for the reference's ``FFNLayer`` (``interface/modules.py:134-174``, post-norm path): ReLU in ``linear1``'s GEMM epilogue, the
    residual accumulated in place by ``linear2``'s, then the LayerNorm."""

    def __init__(self, d_model, dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        if normalize_before or activation != "relu":
            raise NotImplementedError("the B200 FFNLayer implements the post-norm ReLU configuration (step1.yaml)")
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        self.precision = "bf16"
        self._sig = None
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, tgt):
        if not tgt.is_cuda:
            raise RuntimeError("FFNLayer (B200) has no CPU path: the input must be a CUDA tensor")
        if torch.is_grad_enabled() and (tgt.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("FFNLayer (B200) implements the forward pass only: call it under torch.no_grad()")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        mode, adt = (cabi.MODE_BF16, torch.bfloat16) if self.precision == "bf16" else (cabi.MODE_FP32, torch.float32)
        dev = tgt.device
        lib, st = cabi.lib(), cabi.stream_ptr
        shape = tgt.shape
        C = shape[-1]
        with torch.cuda.device(dev):
            ps = list(self.parameters())
            sig = (str(dev), adt) + tuple((p.data_ptr(), p._version) for p in ps)
            if sig != self._sig:
                f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()      # noqa: E731
                self._w1, self._b1 = f(self.linear1.weight).to(adt).contiguous(), f(self.linear1.bias)
                self._w2, self._b2 = f(self.linear2.weight).to(adt).contiguous(), f(self.linear2.bias)
                self._nw, self._nb = f(self.norm.weight), f(self.norm.bias)
                self._sig = sig
            x = tgt.detach().to(torch.float32).contiguous().view(-1, C).clone()
            rows = x.shape[0]
            a = x
            if adt != torch.float32:
                a = torch.empty(rows, C, dtype=adt, device=dev)
                cabi.check(lib.svb_add_cast(x.data_ptr(), None, a.data_ptr(), _odt(adt), x.numel(), st()), "svb_add_cast")
            hid = MaskPredictionHead._linear(mode, a, self._w1, self._b1, torch.empty(rows, self._w1.shape[0], dtype=adt, device=dev), act=2)   # :160
            cabi.check(lib.svb_linear(mode, hid.data_ptr(), hid.stride(0), self._w2.data_ptr(), self._w2.stride(0), rows, C, hid.shape[1],
                                      self._b2.data_ptr(), 0, x.data_ptr(), C, 0, x.data_ptr(), cabi.DTYPE_F32, C, None, 0, 0, 0, st()), "svb_linear")   # :160-161
            out = torch.empty(rows, C, dtype=torch.float32, device=dev)
            cabi.check(lib.svb_layernorm(x.data_ptr(), None, self._nw.data_ptr(), self._nb.data_ptr(), out.data_ptr(), cabi.DTYPE_F32, rows, C,
                                         float(self.norm.eps), st()), "svb_layernorm")                                                        # :162
        return out.view(shape).to(tgt.dtype)


class XDecoderMaskPath(nn.Module):
    """The mask path of ``XDecoder.forward`` (``interface/xdecoder.py:191-329``) for segmentation inference (``task='seg'``, eval mode, no
    grounding / caption tokens): level prompting of the three feature maps, the learnable queries, then per layer masked cross-attention ->
    self-attention -> FFN -> mask branch of the prediction heads, whose attention mask feeds the next layer.  Parameters carry the names
    they have inside ``XDecoder`` (``query_feat``, ``query_embed``, ``level_embed``, ``transformer_{cross,self}_attention_layers.N``,
    ``transformer_ffn_layers.N``, ``decoder_norm``, ``mask_embed``), so those entries of its ``state_dict`` load with ``strict=True``.
    Returns ``{"pred_masks": (B, Q, H, W), "aux_masks": [...]}`` and — when built with ``dim_proj`` / ``bbox`` / ``caption`` —
    ``pred_logits`` (class logits against the text embeddings passed to forward: the language encoder itself is out of scope),
    ``pred_boxes``, ``pred_captions`` and the per-layer ``aux_outputs`` as the reference returns them (:319-327).

    ``in_channels`` / ``enforce_input_project``: the reference inserts a 1x1 ``input_proj`` convolution per level when the pixel
    decoder's width differs from ``hidden_dim`` or ENFORCE_INPUT_PROJ is set (:120-127); step1.yaml uses 512 / 512 / False, which
    makes them empty ``nn.Sequential``s — the only configuration implemented here."""

    def __init__(self, hidden_dim=512, mask_dim=512, num_queries=101, nheads=8, dim_feedforward=2048, num_levels=3,
                 level_indexes=(0, 1, 2, 0, 1, 2, 0, 1, 2), dim_proj=None, bbox=False, caption=False, in_channels=None,
                 enforce_input_project=False):
        super().__init__()
        if enforce_input_project or (in_channels is not None and in_channels != hidden_dim):
            raise NotImplementedError("XDecoderMaskPath (B200): input_proj convolutions (in_channels != hidden_dim or "
                                      "ENFORCE_INPUT_PROJ, xdecoder.py:120-127) are not implemented; step1.yaml does not use them")
        self.num_queries, self.num_heads, self.num_feature_levels = num_queries, nheads, num_levels
        self.level_indexes = list(level_indexes)
        self.num_layers = len(self.level_indexes)
        self.fuse_mask_clear = os.environ.get("SVB_MASK_CLEAR_FUSE", "1") != "0"      # (0: separate svb_mask_clear_full_rows pass per layer, A/B)
        self.transformer_self_attention_layers = nn.ModuleList(SelfAttentionLayer(hidden_dim, nheads) for _ in range(self.num_layers))
        self.transformer_cross_attention_layers = nn.ModuleList(CrossAttentionLayer(hidden_dim, nheads) for _ in range(self.num_layers))
        self.transformer_ffn_layers = nn.ModuleList(FFNLayer(hidden_dim, dim_feedforward) for _ in range(self.num_layers))
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.level_embed = nn.Embedding(num_levels, hidden_dim)
        self._head = [MaskPredictionHead(hidden_dim, mask_dim, num_queries, nheads, dim_proj=dim_proj, bbox=bbox, caption=caption)]
        # parameters registered under XDecoder's names
        self.decoder_norm = self._head[0].decoder_norm
        self.mask_embed = self._head[0].mask_embed
        if self._head[0].class_embed is not None:
            self.class_embed = self._head[0].class_embed
        if self._head[0].bbox_embed is not None:
            self.bbox_embed = self._head[0].bbox_embed
        m = torch.zeros(1, num_queries, num_queries, dtype=torch.bool)                      # xdecoder.py:149-153 (object / class queries)
        m[:, :num_queries - 1, num_queries - 1:] = True
        m[:, num_queries - 1:, :num_queries - 1] = True
        # The reference registers its (1, Q + contxt_len, Q + contxt_len) mask persistently (:153); this slice has no caption tokens, so
        # the buffer is (1, Q, Q), derived from num_queries and not saved.  A reference state_dict's `self_attn_mask` entry is accepted
        # and ignored on load (see _load_from_state_dict), so strict=True loading of the reference's entries works.
        self.register_buffer("self_attn_mask", m, persistent=False)
        self.num_pos_feats, self.temperature, self.scale = hidden_dim // 2, 10000, 2 * 3.141592653589793
        self._pos_cache = {}

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        state_dict.pop(prefix + "self_attn_mask", None)          # the reference's persistent buffer: rebuilt from num_queries here
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    @property
    def precision(self):
        return self._head[0].precision

    @precision.setter
    def precision(self, p):
        self._head[0].precision = p
        for ml in (self.transformer_self_attention_layers, self.transformer_cross_attention_layers, self.transformer_ffn_layers):
            for layer in ml:
                layer.precision = p

    def _pos(self, h, w, bs, device):
        """PositionEmbeddingSine(hidden / 2, normalize=True) of an (h, w) map as (h * w, bs, C) (xdecoder.py:204,208), cached per shape."""
        key = (h, w, bs)
        if key not in self._pos_cache:
            npf = self.num_pos_feats
            dim_t = torch.arange(npf, dtype=torch.float32, device=device)
            dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
            y = torch.arange(1, h + 1, dtype=torch.float32, device=device) / (h + 1e-6) * self.scale
            x = torch.arange(1, w + 1, dtype=torch.float32, device=device) / (w + 1e-6) * self.scale
            py, px = y[:, None] / dim_t, x[:, None] / dim_t
            py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
            px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
            pos = torch.cat((py[:, None, :].expand(h, w, npf), px[None, :, :].expand(h, w, npf)), dim=2).reshape(h * w, 1, 2 * npf)
            self._pos_cache[key] = pos.expand(h * w, bs, 2 * npf).contiguous()
        return self._pos_cache[key]

    def forward(self, x, mask_features, mask_rows=None, mask_shape=None, text_embeddings=None, logit_scale=None):
        """x: three (B, C, H_i, W_i) maps (the pixel decoder's multi_scale_features), mask_features (B, Cm, H, W) — or None with
        `mask_rows` / `mask_shape` from `MSDeformAttnPixelDecoder.forward(..., rows_out=True)`.  `text_embeddings` (K, dim_proj) /
        `logit_scale`: what the language encoder holds as `default_text_embeddings` / `logit_scale` (vlpencoder.py:239-245)."""
        if len(x) != self.num_feature_levels:
            raise AssertionError("x must hold num_feature_levels maps")                    # :195
        if mask_features is None:
            if mask_rows is None or mask_shape is None:
                raise ValueError("mask_features is None: mask_rows and mask_shape are required")
            mask_features = tuple(mask_shape)
        if not x[0].is_cuda:
            raise RuntimeError("XDecoderMaskPath (B200) has no CPU path: the inputs must be CUDA tensors")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("XDecoderMaskPath (B200) implements the forward pass only: call it under torch.no_grad()")
        dev = x[0].device
        lib, st = cabi.lib(), cabi.stream_ptr
        head = self._head[0]
        with torch.cuda.device(dev):
            bs = x[0].shape[0]
            src, pos, size_list = [], [], []
            for i in range(self.num_feature_levels):                                         # :202-209
                h, w = int(x[i].shape[-2]), int(x[i].shape[-1])
                size_list.append((h, w))
                pos.append(self._pos(h, w, bs, dev))
                xi = x[i].detach().contiguous()
                if xi.dtype not in (torch.float32, torch.bfloat16):
                    xi = xi.float()
                seq = torch.empty(h * w, bs, xi.shape[1], dtype=torch.float32, device=dev)
                lvl_e = self.level_embed.weight[i].detach().to(device=dev, dtype=torch.float32).contiguous()
                cabi.check(lib.svb_nchw_to_seq(xi.data_ptr(), _odt(xi.dtype), seq.data_ptr(), cabi.DTYPE_F32, bs, xi.shape[1], h * w,
                                               lvl_e.data_ptr(), st()), "svb_nchw_to_seq")          # flatten + level embedding + permute(2, 0, 1)
                src.append(seq)
            query_embed = self.query_embed.weight.detach().float().unsqueeze(1).repeat(1, bs, 1)      # :214-215
            output = self.query_feat.weight.detach().float().unsqueeze(1).repeat(1, bs, 1)
            self_mask = self.self_attn_mask.to(dev).repeat(bs * self.num_heads, 1, 1).contiguous()    # :254
            masks, extras = [], []
            mrows = mask_rows if mask_rows is not None else head.mask_rows(mask_features)            # once for the ten prediction-head calls
            # the clearing of fully masked rows (:267, first statement of every layer) is applied while the mask is written
            kw = dict(rows=mrows, text_embeddings=text_embeddings, logit_scale=logit_scale, clear_full_rows=self.fuse_mask_clear)
            res = head(output, mask_features, size_list[0], **kw)                                    # :257
            masks.append(res["outputs_mask"])
            extras.append(res)
            attn_mask = res["attn_mask"]
            level_ops = {}                                                                           # operand copies of src[lvl] (+ pos[lvl]), per level
            for i in range(self.num_layers):
                lvl = self.level_indexes[i]
                if not self.fuse_mask_clear:
                    cabi.check(lib.svb_mask_clear_full_rows(attn_mask.data_ptr(), attn_mask.shape[0] * attn_mask.shape[1], attn_mask.shape[2], st()),
                               "svb_mask_clear_full_rows")                                            # :267
                output, _ = self.transformer_cross_attention_layers[i](output, src[lvl], memory_mask=attn_mask, pos=pos[lvl],
                                                                       query_pos=query_embed,
                                                                       memory_operands=level_ops.setdefault(lvl, {}))    # :272-277
                output = self.transformer_self_attention_layers[i](output, tgt_mask=self_mask, query_pos=query_embed)   # :283-287
                output = self.transformer_ffn_layers[i](output)                                       # :290-292
                res = head(output, mask_features, size_list[(i + 1) % self.num_feature_levels], **kw)         # :299
                attn_mask = res["attn_mask"]
                masks.append(res["outputs_mask"])
                extras.append(res)
        out = {"pred_masks": masks[-1], "aux_masks": masks[:-1]}
        if head.class_embed is not None or head.bbox_embed is not None:                               # :319-327
            out.update(pred_logits=extras[-1]["outputs_class"], pred_boxes=extras[-1]["outputs_bbox"],
                       pred_captions=extras[-1]["outputs_caption"],
                       aux_outputs=[{"pred_logits": e["outputs_class"], "pred_masks": m, "pred_boxes": e["outputs_bbox"],
                                     "pred_captions": e["outputs_caption"]} for e, m in zip(extras[:-1], masks[:-1])])
        return out
