"""CPU ORACLE for the mask branch of the X-Decoder prediction heads (scope row N4, first slice).  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` and tools that check the CUDA path may import this file; the product never does.

``mask_branch`` restates ``XDecoder.forward_prediction_heads`` (``/root/reference/modeling/interface/xdecoder.py:429-470``) for the
inference path with the mask task on, on a plain state_dict with the reference's keys (``decoder_norm.*``, ``mask_embed.layers.N.*``);
``resize_bicubic_aa`` writes out what ``F.interpolate(mode="bicubic", align_corners=False, antialias=True)`` (:463) computes — the
separable antialiased cubic-convolution filter (a = -0.5, support 2 * max(scale, 1), weights normalised per output pixel) — with
explicit weight matrices.

Parity pin: ``tests/golden/mask_head_*.npz`` = outputs of the UNMODIFIED reference method executed on CPU
(``tests/golden/make_golden_mask_head.py``); ``tests/test_oracle.py`` checks this file against them and ``resize_bicubic_aa``
against ``F.interpolate`` itself.
"""
from __future__ import annotations

import torch


def _cubic(x, a=-0.5):
    x = x.abs()
    return torch.where(x < 1, ((a + 2) * x - (a + 3)) * x * x + 1, torch.where(x < 2, (((x - 5) * x + 8) * x - 4) * a, torch.zeros_like(x)))


def aa_weights(n_in, n_out, dtype=torch.float64):
    """(n_out, n_in) matrix of the antialiased bicubic filter along one axis."""
    scale = n_in / n_out
    support = 2.0 * scale if scale >= 1 else 2.0
    inv = 1.0 / scale if scale >= 1 else 1.0
    wm = torch.zeros(n_out, n_in, dtype=dtype)
    for i in range(n_out):
        center = scale * (i + 0.5)
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), n_in)
        j = torch.arange(lo, hi, dtype=dtype)
        w = _cubic((j - center + 0.5) * inv)
        wm[i, lo:hi] = w / w.sum()
    return wm


def resize_bicubic_aa(x, size):
    """x (..., H, W) -> (..., size[0], size[1])."""
    wh = aa_weights(x.shape[-2], size[0], x.dtype)
    ww = aa_weights(x.shape[-1], size[1], x.dtype)
    return torch.einsum("oh,...hw,pw->...op", wh, x, ww)


def mask_branch(sd, output, mask_features, attn_mask_target_size, num_queries, num_heads):
    """-> (outputs_mask (B, Q, H, W), attention logits (B, Q, h, w), attn_mask (B * heads, Q, h * w) bool)."""
    c = output.shape[-1]
    mu = output.mean(-1, keepdim=True)
    var = ((output - mu) ** 2).mean(-1, keepdim=True)
    dec = ((output - mu) / torch.sqrt(var + 1e-5) * sd["decoder_norm.weight"] + sd["decoder_norm.bias"]).transpose(0, 1)       # :430-431
    nrm = dec / (dec.norm(dim=-1, keepdim=True) + 1e-7)                                                                          # :440
    obj, cls = nrm[:, :num_queries - 1], nrm[:, num_queries - 1:num_queries]
    sim = (cls @ obj.transpose(1, 2)).softmax(-1)[:, 0, :, None]                                                                 # :444
    cls_token = (sim * dec[:, :num_queries - 1]).sum(dim=1, keepdim=True)                                                        # :445
    dec = torch.cat((dec[:, :num_queries - 1], cls_token), dim=1)                                                                # :450
    x = dec
    n = len([k for k in sd if k.startswith("mask_embed.layers.") and k.endswith(".weight")])
    for i in range(n):                                                                                                           # :458
        x = x @ sd[f"mask_embed.layers.{i}.weight"].t() + sd[f"mask_embed.layers.{i}.bias"]
        if i < n - 1:
            x = torch.relu(x)
    outputs_mask = torch.einsum("bqc,bchw->bqhw", x, mask_features)                                                              # :459
    logits = resize_bicubic_aa(outputs_mask, attn_mask_target_size)                                                              # :463
    attn = (logits.sigmoid().flatten(2).unsqueeze(1).repeat(1, num_heads, 1, 1).flatten(0, 1) < 0.5)                             # :467
    return outputs_mask, logits, attn


def decoder_output_rows(sd, output, num_queries):
    """decoder_norm + transpose + class-token recompute (xdecoder.py:430-450): (Q, B, C) -> (B, Q, C)."""
    mu = output.mean(-1, keepdim=True)
    var = ((output - mu) ** 2).mean(-1, keepdim=True)
    dec = ((output - mu) / torch.sqrt(var + 1e-5) * sd["decoder_norm.weight"] + sd["decoder_norm.bias"]).transpose(0, 1)
    nrm = dec / (dec.norm(dim=-1, keepdim=True) + 1e-7)
    obj, cls = nrm[:, :num_queries - 1], nrm[:, num_queries - 1:num_queries]
    sim = (cls @ obj.transpose(1, 2)).softmax(-1)[:, 0, :, None]
    cls_token = (sim * dec[:, :num_queries - 1]).sum(dim=1, keepdim=True)
    return torch.cat((dec[:, :num_queries - 1], cls_token), dim=1)


def class_box_branch(sd, output, num_queries, text_embeddings, logit_scale):
    """The remaining outputs of forward_prediction_heads (xdecoder.py:452-484) for the inference path:
    class_embed = decoder_output @ class_embed (:453); outputs_class = LanguageEncoder.compute_similarity(class_embed)
    (modeling/language/vlpencoder.py:239-245: exp(logit_scale) * normalise(v) @ t_emb^T, normalisation with + 1e-7);
    outputs_bbox = bbox_embed(decoder_output) (a 3-layer ReLU MLP, :478); outputs_caption = class_embed (:482).
    -> (outputs_class (B, Q, K), outputs_bbox (B, Q, 4), outputs_caption (B, Q, dim_proj)).  Pinned by tests/golden/mask_head_full_*.npz."""
    dec = decoder_output_rows(sd, output, num_queries)
    class_embed = dec @ sd["class_embed"]
    v = class_embed / (class_embed.norm(dim=-1, keepdim=True) + 1e-7)
    outputs_class = torch.as_tensor(logit_scale).exp() * v @ text_embeddings.unsqueeze(0).transpose(1, 2)
    x = dec
    n = len([k for k in sd if k.startswith("bbox_embed.layers.") and k.endswith(".weight")])
    for i in range(n):
        x = x @ sd[f"bbox_embed.layers.{i}.weight"].t() + sd[f"bbox_embed.layers.{i}.bias"]
        if i < n - 1:
            x = torch.relu(x)
    return outputs_class, x, class_embed


def cross_attention_layer(sd, tgt, memory, memory_mask, pos, query_pos, n_heads, mha="multihead_attn"):
    """CrossAttentionLayer.forward_post (interface/modules.py:95-106) with torch's nn.MultiheadAttention written out (in_proj split in
    q / k / v, q scaled by head_dim^-0.5, additive -inf mask where memory_mask is True, softmax over the keys, out_proj), eval mode.
    tgt (Q, B, C), memory (HW, B, C) sequence-first; state_dict keys as the reference's.  Pinned by tests/golden/cross_attn_*.npz."""
    Q, B, C = tgt.shape
    HW = memory.shape[0]
    d = C // n_heads
    w, b = sd[mha + ".in_proj_weight"], sd[mha + ".in_proj_bias"]
    qi = tgt if query_pos is None else tgt + query_pos
    ki = memory if pos is None else memory + pos
    q = qi @ w[:C].t() + b[:C]
    k = ki @ w[C:2 * C].t() + b[C:2 * C]
    v = memory @ w[2 * C:].t() + b[2 * C:]
    q = q.reshape(Q, B * n_heads, d).transpose(0, 1) * (d ** -0.5)
    k = k.reshape(HW, B * n_heads, d).transpose(0, 1)
    v = v.reshape(HW, B * n_heads, d).transpose(0, 1)
    s = q @ k.transpose(1, 2)
    if memory_mask is not None:
        s = s.masked_fill(memory_mask, float("-inf"))
    a = torch.softmax(s, -1) @ v                                           # (B * heads, Q, d)
    a = a.transpose(0, 1).reshape(Q, B, C)
    tgt2 = a @ sd[mha + ".out_proj.weight"].t() + sd[mha + ".out_proj.bias"]
    x = tgt + tgt2
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * sd["norm.weight"] + sd["norm.bias"]


def self_attention_layer(sd, tgt, tgt_mask, query_pos, n_heads):
    """SelfAttentionLayer.forward_post (interface/modules.py:37-47): q = k = tgt + query_pos, value = tgt."""
    return cross_attention_layer(sd, tgt, tgt, tgt_mask, query_pos, query_pos, n_heads, mha="self_attn")


def ffn_layer(sd, tgt):
    """FFNLayer.forward_post (interface/modules.py:159-163)."""
    x = tgt + torch.relu(tgt @ sd["linear1.weight"].t() + sd["linear1.bias"]) @ sd["linear2.weight"].t() + sd["linear2.bias"]
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * sd["norm.weight"] + sd["norm.bias"]


def xdecoder_mask_path(sd, x, mask_features, num_queries, num_heads, level_indexes):
    """XDecoder.forward for task='seg' in eval mode without grounding / caption tokens (interface/xdecoder.py:191-329), mask outputs only:
    -> list of outputs_mask, one per prediction-head call (the last one is `pred_masks`)."""
    from . import pixel_decoder_oracle as po
    hidden = sd["query_feat.weight"].shape[1]
    src, pos, sizes = [], [], []
    for i, xi in enumerate(x):                                                                           # :202-209
        sizes.append(tuple(xi.shape[-2:]))
        pos.append(po.position_embedding_sine(xi, hidden // 2).flatten(2).permute(2, 0, 1))
        src.append((xi.flatten(2) + sd["level_embed.weight"][i][None, :, None]).permute(2, 0, 1))
    bs = src[0].shape[1]
    query_embed = sd["query_embed.weight"].unsqueeze(1).repeat(1, bs, 1)                                 # :214-215
    output = sd["query_feat.weight"].unsqueeze(1).repeat(1, bs, 1)
    q = num_queries
    sam = torch.zeros(1, q, q, dtype=torch.bool)                                                         # :149-153, :254
    sam[:, :q - 1, q - 1:] = True
    sam[:, q - 1:, :q - 1] = True
    sam = sam.repeat(bs * num_heads, 1, 1)
    head = {k: v for k, v in sd.items() if k.startswith(("decoder_norm.", "mask_embed."))}
    masks = []
    m, _, attn = mask_branch(head, output, mask_features, sizes[0], q, num_heads)                        # :257
    masks.append(m)
    for i, lvl in enumerate(level_indexes):
        attn = attn.clone()
        attn[torch.where(attn.sum(-1) == attn.shape[-1])] = False                                         # :267
        sub = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}                  # noqa: E731
        output = cross_attention_layer(sub(f"transformer_cross_attention_layers.{i}."), output, src[lvl], attn, pos[lvl], query_embed, num_heads)
        output = self_attention_layer(sub(f"transformer_self_attention_layers.{i}."), output, sam, query_embed, num_heads)
        output = ffn_layer(sub(f"transformer_ffn_layers.{i}."), output)
        m, _, attn = mask_branch(head, output, mask_features, sizes[(i + 1) % len(x)], q, num_heads)     # :299
        masks.append(m)
    return masks
