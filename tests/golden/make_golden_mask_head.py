"""Golden fixtures of the mask branch of the X-Decoder prediction heads, FROM THE UNMODIFIED REFERENCE METHOD.

Run in the build container only:  python tests/golden/make_golden_mask_head.py

``modeling/interface/xdecoder.py`` imports timm / detectron2 / fvcore at module level (absent here).  The SOURCE TEXT of
``XDecoder.forward_prediction_heads`` (:429-494) is cut out with ``ast`` and executed unmodified as a plain function on a stand-in
``self`` that carries exactly the attributes the method reads: ``decoder_norm`` (nn.LayerNorm), ``mask_embed`` (the reference's own
``MLP`` class, source text of ``interface/modules.py:188-201``), ``class_embed``, ``num_queries``, ``num_heads``, ``task_switch`` (mask
on, everything else off), ``training = False`` and a ``lang_encoder`` whose ``compute_similarity`` returns None (the class logits
are not part of this slice).  Runs in fp32 on the CPU.
"""
import ast
import os
import types

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/modeling/interface"


def _cut(path, cls, fn=None):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            if fn is None:
                return node
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == fn:
                    return sub
    raise KeyError((cls, fn))


ns = {"torch": torch, "nn": nn, "F": F}
exec(compile(ast.Module(body=[_cut(REF + "/modules.py", "MLP")], type_ignores=[]), "modules.py", "exec"), ns)
exec(compile(ast.Module(body=[_cut(REF + "/xdecoder.py", "XDecoder", "forward_prediction_heads")], type_ignores=[]), "xdecoder.py", "exec"), ns)
RefMLP, ref_heads = ns["MLP"], ns["forward_prediction_heads"]

# name -> (hidden, mask_dim, queries, heads, batch, mask side, attention-mask size)
CASES = {
    "small": (64, 32, 11, 4, 2, (32, 32), (8, 8)),
    "odd": (128, 64, 101, 8, 1, (48, 40), (12, 10)),        # the reference's 101 queries; a 4x reduction on a non-square map
    "half": (64, 64, 7, 2, 1, (24, 24), (12, 12)),          # scale 2 (the 128^2 level of a 256^2 mask)
}
for seed, (name, (C, MD, Q, NH, B, (H, W), tgt)) in enumerate(CASES.items()):
    g = torch.Generator().manual_seed(1700 + seed)
    torch.manual_seed(1800 + seed)
    me = types.SimpleNamespace()
    me.decoder_norm = nn.LayerNorm(C)
    me.mask_embed = RefMLP(C, C, MD, 3)
    with torch.no_grad():
        me.decoder_norm.weight.add_(torch.randn(C, generator=g) * 0.2)
        me.decoder_norm.bias.add_(torch.randn(C, generator=g) * 0.2)
    me.class_embed = torch.randn(C, 16, generator=g)
    me.num_queries, me.num_heads, me.training = Q, NH, False
    me.task_switch = {"mask": True, "bbox": False, "caption": False, "captioning": False, "grounding": False}
    me.lang_encoder = types.SimpleNamespace(compute_similarity=lambda x, fake=False: None)
    output = torch.randn(Q, B, C, generator=g)
    mask_features = torch.randn(B, MD, H, W, generator=g)
    with torch.no_grad():
        res = ref_heads(me, output, mask_features, attn_mask_target_size=tgt)
    blob = {"output": output.numpy(), "mask_features": mask_features.numpy(), "outputs_mask": res["outputs_mask"].numpy(),
            "attn_mask": res["attn_mask"].numpy(), "meta": np.array([C, MD, Q, NH, tgt[0], tgt[1]], dtype=np.int64)}
    blob["sd.decoder_norm.weight"] = me.decoder_norm.weight.detach().numpy()
    blob["sd.decoder_norm.bias"] = me.decoder_norm.bias.detach().numpy()
    for k, v in me.mask_embed.state_dict().items():
        blob["sd.mask_embed." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"mask_head_{name}.npz"), **blob)
    print(name, tuple(res["outputs_mask"].shape), tuple(res["attn_mask"].shape), float(res["attn_mask"].float().mean()))


# ---- the remaining outputs of forward_prediction_heads (:452-484): class logits (class_embed projection + the language encoder's
# compute_similarity, modeling/language/vlpencoder.py:239-245, executed from ITS source text on a stand-in that carries logit_scale
# and the text embeddings), box MLP, caption embeddings.  task_switch: mask, bbox, caption on. ----
exec(compile(ast.Module(body=[_cut("/root/reference/modeling/language/vlpencoder.py", "LanguageEncoder", "compute_similarity")], type_ignores=[]),
             "vlpencoder.py", "exec"), ns)
ref_similarity = ns["compute_similarity"]
FULL = {
    "full_small": (64, 32, 11, 4, 2, (32, 32), (8, 8), 48, 13),       # ..., dim_proj, number of text embeddings (classes)
    "full_q101": (128, 64, 101, 8, 2, (32, 32), (8, 8), 96, 134),     # 133 COCO classes + background
}
for seed, (name, (C, MD, Q, NH, B, (H, W), tgt, DP, NC)) in enumerate(FULL.items()):
    g = torch.Generator().manual_seed(2700 + seed)
    torch.manual_seed(2800 + seed)
    me = types.SimpleNamespace()
    me.decoder_norm = nn.LayerNorm(C)
    me.mask_embed = RefMLP(C, C, MD, 3)
    me.bbox_embed = RefMLP(C, C, 4, 3)
    with torch.no_grad():
        me.decoder_norm.weight.add_(torch.randn(C, generator=g) * 0.2)
        me.decoder_norm.bias.add_(torch.randn(C, generator=g) * 0.2)
    me.class_embed = torch.randn(C, DP, generator=g) * 0.05
    me.num_queries, me.num_heads, me.training = Q, NH, False
    me.task_switch = {"mask": True, "bbox": True, "caption": True, "captioning": False, "grounding": False}
    lang = types.SimpleNamespace()
    lang.logit_scale = torch.tensor(2.3)                                       # exp = 10
    t_emb = torch.randn(NC, DP, generator=g)
    lang.default_text_embeddings = t_emb / t_emb.norm(dim=-1, keepdim=True)    # the encoder stores normalised text embeddings
    lang.compute_similarity = types.MethodType(ref_similarity, lang)
    me.lang_encoder = lang
    output = torch.randn(Q, B, C, generator=g)
    mask_features = torch.randn(B, MD, H, W, generator=g)
    with torch.no_grad():
        res = ref_heads(me, output, mask_features, attn_mask_target_size=tgt)
    blob = {"output": output.numpy(), "mask_features": mask_features.numpy(), "outputs_mask": res["outputs_mask"].numpy(),
            "attn_mask": res["attn_mask"].numpy(), "outputs_class": res["outputs_class"].numpy(), "outputs_bbox": res["outputs_bbox"].numpy(),
            "outputs_caption": res["outputs_caption"].numpy(), "text_embeddings": lang.default_text_embeddings.numpy(),
            "logit_scale": lang.logit_scale.numpy(), "meta": np.array([C, MD, Q, NH, tgt[0], tgt[1], DP, NC], dtype=np.int64)}
    blob["sd.decoder_norm.weight"] = me.decoder_norm.weight.detach().numpy()
    blob["sd.decoder_norm.bias"] = me.decoder_norm.bias.detach().numpy()
    blob["sd.class_embed"] = me.class_embed.numpy()
    for k, v in me.mask_embed.state_dict().items():
        blob["sd.mask_embed." + k] = v.numpy()
    for k, v in me.bbox_embed.state_dict().items():
        blob["sd.bbox_embed." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"mask_head_{name}.npz"), **blob)
    print(name, tuple(res["outputs_class"].shape), tuple(res["outputs_bbox"].shape), tuple(res["outputs_caption"].shape))
