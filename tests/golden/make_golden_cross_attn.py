"""Golden fixtures of the X-Decoder's masked cross-attention layer, FROM THE UNMODIFIED REFERENCE CLASS.

Run in the build container only:  python tests/golden/make_golden_cross_attn.py

``modeling/interface/modules.py`` imports timm / detectron2 / fvcore at module level (absent here); the SOURCE TEXT of
``CrossAttentionLayer`` (:72-131) and ``_get_activation_fn`` is cut out with ``ast`` and executed unmodified (it only needs torch:
the attention itself is ``nn.MultiheadAttention``).  fp32 on the CPU, eval mode.  Masks as the decoder builds them
(xdecoder.py:467-470): one boolean map per (image, head), True = not allowed, no fully-masked row (the decoder clears those, :260).
"""
import ast
import os

import numpy as np
import torch
from torch import nn, Tensor
from torch.nn import functional as F
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
tree = ast.parse(open("/root/reference/modeling/interface/modules.py").read())
ns = {"torch": torch, "nn": nn, "F": F, "Tensor": Tensor, "Optional": Optional}
for node in tree.body:
    if (isinstance(node, ast.ClassDef) and node.name == "CrossAttentionLayer") or (isinstance(node, ast.FunctionDef) and node.name == "_get_activation_fn"):
        exec(compile(ast.Module(body=[node], type_ignores=[]), "modules.py", "exec"), ns)
RefLayer = ns["CrossAttentionLayer"]

# name -> (d_model, heads, queries, batch, keys (h, w), masked fraction)
CASES = {"small": (128, 2, 11, 2, (8, 8), 0.5), "q101": (128, 2, 101, 1, (24, 20), 0.6), "nomask": (64, 1, 7, 2, (5, 9), None)}
for seed, (name, (C, NH, Q, B, (h, w), frac)) in enumerate(CASES.items()):
    g = torch.Generator().manual_seed(2100 + seed)
    torch.manual_seed(2200 + seed)
    layer = RefLayer(C, NH, dropout=0.0, normalize_before=False).eval()
    with torch.no_grad():
        layer.multihead_attn.in_proj_bias.copy_(torch.randn(3 * C, generator=g) * 0.2)
        layer.multihead_attn.out_proj.bias.copy_(torch.randn(C, generator=g) * 0.2)
        layer.norm.weight.add_(torch.randn(C, generator=g) * 0.2)
        layer.norm.bias.add_(torch.randn(C, generator=g) * 0.2)
    HW = h * w
    tgt, memory = torch.randn(Q, B, C, generator=g), torch.randn(HW, B, C, generator=g)
    pos, query_pos = torch.randn(HW, B, C, generator=g), torch.randn(Q, B, C, generator=g)
    mask = None
    if frac is not None:
        mask = (torch.rand(B, 1, Q, HW, generator=g) < frac).repeat(1, NH, 1, 1).flatten(0, 1)
        mask[torch.where(mask.sum(-1) == mask.shape[-1])] = False
    with torch.no_grad():
        out, _ = layer(tgt, memory, memory_mask=mask, memory_key_padding_mask=None, pos=pos, query_pos=query_pos)
    blob = {"tgt": tgt.numpy(), "memory": memory.numpy(), "pos": pos.numpy(), "query_pos": query_pos.numpy(), "out": out.numpy(),
            "mask": mask.numpy() if mask is not None else np.zeros(0, dtype=bool), "meta": np.array([C, NH], dtype=np.int64)}
    for k, v in layer.state_dict().items():
        blob["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"cross_attn_{name}.npz"), **blob)
    print(name, tuple(out.shape), float(out.abs().mean()))
