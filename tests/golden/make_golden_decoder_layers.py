"""Golden fixtures of the X-Decoder's self-attention and FFN layers, FROM THE UNMODIFIED REFERENCE CLASSES.

Run in the build container only:  python tests/golden/make_golden_decoder_layers.py

As make_golden_cross_attn.py: the SOURCE TEXT of ``SelfAttentionLayer`` (interface/modules.py:14-69) and ``FFNLayer`` (:134-174) is cut
out with ``ast`` and executed unmodified; ``MultiheadAttention`` is the reference's own copy (``modeling/utils/attention.py``, imported
as a file: it needs torch only).  The self-attention mask is the decoder's (xdecoder.py:149-153, 269): object queries and the class
query do not attend to each other.  fp32 on the CPU, eval mode.
"""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch
from torch import nn, Tensor
from torch.nn import functional as F
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_attention", "/root/reference/modeling/utils/attention.py")
att = importlib.util.module_from_spec(spec)
sys.modules["ref_attention"] = att
spec.loader.exec_module(att)
tree = ast.parse(open("/root/reference/modeling/interface/modules.py").read())
ns = {"torch": torch, "nn": nn, "F": F, "Tensor": Tensor, "Optional": Optional, "MultiheadAttention": att.MultiheadAttention}
for node in tree.body:
    if (isinstance(node, ast.ClassDef) and node.name in ("SelfAttentionLayer", "FFNLayer")) or (isinstance(node, ast.FunctionDef) and node.name == "_get_activation_fn"):
        exec(compile(ast.Module(body=[node], type_ignores=[]), "modules.py", "exec"), ns)

# name -> (d_model, heads, queries, batch, d_ffn)
CASES = {"small": (128, 2, 11, 2, 256), "q101": (128, 2, 101, 2, 512)}
for seed, (name, (C, NH, Q, B, FF)) in enumerate(CASES.items()):
    g = torch.Generator().manual_seed(2300 + seed)
    torch.manual_seed(2400 + seed)
    sa = ns["SelfAttentionLayer"](C, NH, dropout=0.0, normalize_before=False).eval()
    ffn = ns["FFNLayer"](C, FF, dropout=0.0, normalize_before=False).eval()
    with torch.no_grad():
        sa.self_attn.in_proj_bias.copy_(torch.randn(3 * C, generator=g) * 0.2)
        sa.self_attn.out_proj.bias.copy_(torch.randn(C, generator=g) * 0.2)
        for m in (sa.norm, ffn.norm):
            m.weight.add_(torch.randn(C, generator=g) * 0.2)
            m.bias.add_(torch.randn(C, generator=g) * 0.2)
        ffn.linear1.bias.copy_(torch.randn(FF, generator=g) * 0.2)
        ffn.linear2.bias.copy_(torch.randn(C, generator=g) * 0.2)
    tgt, qpos = torch.randn(Q, B, C, generator=g), torch.randn(Q, B, C, generator=g)
    m1 = torch.zeros(1, Q, Q, dtype=torch.bool)                      # xdecoder.py:149-153 without caption tokens
    m1[:, :Q - 1, Q - 1:Q] = True
    m1[:, Q - 1:Q, :Q - 1] = True
    mask = m1.repeat(B * NH, 1, 1)                                   # :269
    with torch.no_grad():
        y = sa(tgt, tgt_mask=mask, tgt_key_padding_mask=None, query_pos=qpos)
        zf = ffn(y)
    blob = {"tgt": tgt.numpy(), "query_pos": qpos.numpy(), "mask": mask.numpy(), "self_out": y.numpy(), "ffn_out": zf.numpy(),
            "meta": np.array([C, NH, FF], dtype=np.int64)}
    for k, v in sa.state_dict().items():
        blob["sa." + k] = v.numpy()
    for k, v in ffn.state_dict().items():
        blob["ffn." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"decoder_layers_{name}.npz"), **blob)
    print(name, tuple(y.shape), float(y.abs().mean()), float(zf.abs().mean()), sorted(sa.state_dict().keys())[:3])
