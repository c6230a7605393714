"""Golden fixtures of the X-Decoder's mask path (segmentation inference), FROM THE UNMODIFIED REFERENCE METHODS.

Run in the build container only:  python tests/golden/make_golden_xdecoder_mask_path.py

The SOURCE TEXT of ``XDecoder.forward`` (interface/xdecoder.py:191-329), ``forward_prediction_heads`` (:429-494) and ``_set_aux_loss`` is
cut out with ``ast`` and executed unmodified as methods of a stand-in object that carries the attributes those methods read: the
reference's own layer classes (``CrossAttentionLayer`` / ``SelfAttentionLayer`` / ``FFNLayer`` / ``MLP``, source text of
interface/modules.py; ``MultiheadAttention`` = modeling/utils/attention.py), its ``PositionEmbeddingSine``, embeddings, the
``self_attn_mask`` buffer built as at :149-154, ``task_switch`` with only ``mask`` on, ``training = False`` and a ``lang_encoder`` whose
``compute_similarity`` returns None (class logits are not part of this path).  ``task='seg'``; fp32 on the CPU.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch import nn, Tensor
from torch.nn import functional as F
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/modeling"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


att = _load("ref_attention", REF + "/utils/attention.py")
posenc = _load("ref_position_encoding", REF + "/modules/position_encoding.py")
ns = {"torch": torch, "nn": nn, "F": F, "Tensor": Tensor, "Optional": Optional, "MultiheadAttention": att.MultiheadAttention}
for node in ast.parse(open(REF + "/interface/modules.py").read()).body:
    if (isinstance(node, ast.ClassDef) and node.name in ("SelfAttentionLayer", "CrossAttentionLayer", "FFNLayer", "MLP")) or \
            (isinstance(node, ast.FunctionDef) and node.name == "_get_activation_fn"):
        exec(compile(ast.Module(body=[node], type_ignores=[]), "modules.py", "exec"), ns)
methods = {}
for node in ast.parse(open(REF + "/interface/xdecoder.py").read()).body:
    if isinstance(node, ast.ClassDef) and node.name == "XDecoder":
        for sub in node.body:
            if isinstance(sub, ast.FunctionDef) and sub.name in ("forward", "forward_prediction_heads", "_set_aux_loss"):
                exec(compile(ast.Module(body=[sub], type_ignores=[]), "xdecoder.py", "exec"), ns)
                methods[sub.name] = ns[sub.name]

# name -> (hidden, mask_dim, queries, heads, d_ffn, batch, level sides, mask side, layers)
CASES = {"small": (64, 32, 11, 1, 128, 2, (4, 8, 16), 32, 9), "q101": (128, 64, 101, 2, 128, 1, (6, 12, 24), 40, 3)}
for seed, (name, (C, MD, Q, NH, FF, B, sides, ms, NL)) in enumerate(CASES.items()):
    g = torch.Generator().manual_seed(2600 + seed)
    torch.manual_seed(2700 + seed)
    me = types.SimpleNamespace()
    me.num_feature_levels, me.num_queries, me.num_heads, me.contxt_len, me.training = 3, Q, NH, 7, False
    me.level_indexes = [0, 1, 2, 0, 1, 2, 0, 1, 2][:NL]
    me.num_layers = NL
    me.pe_layer = posenc.PositionEmbeddingSine(C // 2, normalize=True)
    me.input_proj = nn.ModuleList(nn.Sequential() for _ in range(3))
    me.level_embed, me.query_feat, me.query_embed = nn.Embedding(3, C), nn.Embedding(Q, C), nn.Embedding(Q, C)
    me.transformer_self_attention_layers = nn.ModuleList(ns["SelfAttentionLayer"](C, NH, dropout=0.0, normalize_before=False) for _ in range(NL))
    me.transformer_cross_attention_layers = nn.ModuleList(ns["CrossAttentionLayer"](C, NH, dropout=0.0, normalize_before=False) for _ in range(NL))
    me.transformer_ffn_layers = nn.ModuleList(ns["FFNLayer"](C, FF, dropout=0.0, normalize_before=False) for _ in range(NL))
    me.decoder_norm, me.mask_embed = nn.LayerNorm(C), ns["MLP"](C, C, MD, 3)
    me.class_embed = torch.randn(C, 16, generator=g)
    me.task_switch = {"mask": True, "bbox": False, "caption": False, "captioning": False, "grounding": False}
    me.lang_encoder = types.SimpleNamespace(compute_similarity=lambda x, fake=False: None)
    me.mask_classification = True
    sam = torch.zeros((1, Q + me.contxt_len, Q + me.contxt_len)).bool()                       # xdecoder.py:149-153
    sam[:, :Q, Q:] = True
    sam[:, Q:, Q:] = torch.triu(torch.ones((1, me.contxt_len, me.contxt_len)), diagonal=1).bool()
    sam[:, :Q - 1, Q - 1:Q] = True
    sam[:, Q - 1:Q, :Q - 1] = True
    me.self_attn_mask = sam
    for k_, fn in methods.items():
        setattr(me, k_, types.MethodType(fn, me))
    mods = {"level_embed": me.level_embed, "query_feat": me.query_feat, "query_embed": me.query_embed, "decoder_norm": me.decoder_norm,
            "mask_embed": me.mask_embed, "transformer_self_attention_layers": me.transformer_self_attention_layers,
            "transformer_cross_attention_layers": me.transformer_cross_attention_layers, "transformer_ffn_layers": me.transformer_ffn_layers}
    holder = nn.ModuleDict(mods).eval()
    with torch.no_grad():
        for k_, p in holder.named_parameters():
            if k_.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif "norm" in k_:
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
    x = [torch.randn(B, C, s, s, generator=g) for s in sides]
    mask_features = torch.randn(B, MD, ms, ms, generator=g)
    with torch.no_grad():
        out = me.forward(x, mask_features, task="seg")
    blob = {"pred_masks": out["pred_masks"].numpy(), "mask_features": mask_features.numpy(),
            "meta": np.array([C, MD, Q, NH, FF, NL], dtype=np.int64)}
    for i, a in enumerate(out["aux_outputs"]):
        if Q < 32:                                       # the per-layer masks only for the small case (fixture size)
            blob[f"aux{i}"] = a["pred_masks"].numpy()
    for i, t in enumerate(x):
        blob[f"x{i}"] = t.numpy()
    for k_, v in holder.state_dict().items():
        blob["sd." + k_] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"xdecoder_mask_path_{name}.npz"), **blob)
    print(name, tuple(out["pred_masks"].shape), len(out["aux_outputs"]), float(out["pred_masks"].abs().mean()))
