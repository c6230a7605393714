"""Golden fixtures of the deformable-attention encoder, FROM THE UNMODIFIED REFERENCE CLASSES.

Run in the build container only:  python tests/golden/make_golden_deform_encoder.py

``transformer_encoder_deform.py`` imports detectron2 / fvcore at module level (absent here), so the module cannot be imported; the
SOURCE TEXT of its three encoder classes (``MSDeformAttnTransformerEncoderOnly``, ``...EncoderLayer``, ``...Encoder``, lines 23-161) is
cut out with ``ast`` and executed unmodified in a namespace that holds what those classes use: torch, the reference's own
``MSDeformAttn`` (imported as in make_golden_msda_module.py, taking its pure-PyTorch branch) and the reference's ``_get_clones`` /
``_get_activation_fn`` (``transformer_blocks.py``, importable).  Modules run in fp64 / eval(); everything is stored as fp32.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ENC = "/root/reference/modeling/vision/encoder"
OPS = ENC + "/ops"

stub = types.ModuleType("MultiScaleDeformableAttention")


def _absent(*a, **k):
    raise RuntimeError("the compiled MultiScaleDeformableAttention extension is absent")


stub.ms_deform_attn_forward = _absent
stub.ms_deform_attn_backward = _absent
sys.modules["MultiScaleDeformableAttention"] = stub


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


pkg = types.ModuleType("refops")
pkg.__path__ = [OPS]
sys.modules["refops"] = pkg
_load("refops.functions", OPS + "/functions/__init__.py")
RefMSDeformAttn = _load("refops.modules", OPS + "/modules/__init__.py").MSDeformAttn
blocks = _load("ref_transformer_blocks", ENC + "/transformer_blocks.py")

text = open(ENC + "/transformer_encoder_deform.py").read()
wanted = ("MSDeformAttnTransformerEncoderOnly", "MSDeformAttnTransformerEncoderLayer", "MSDeformAttnTransformerEncoder")
ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional, "normal_": torch.nn.init.normal_, "MSDeformAttn": RefMSDeformAttn,
      "_get_clones": blocks._get_clones, "_get_activation_fn": blocks._get_activation_fn}
for node in ast.parse(text).body:
    if isinstance(node, ast.ClassDef) and node.name in wanted:
        exec(compile(ast.Module(body=[node], type_ignores=[]), "transformer_encoder_deform.py", "exec"), ns)
RefEncoderOnly = ns["MSDeformAttnTransformerEncoderOnly"]

# name -> (d_model, heads, layers, d_ffn, points, level shapes, batch)
CASES = {
    "small": (64, 4, 2, 128, 4, [(8, 8), (4, 4), (2, 2)], 2),
    "heads64": (256, 4, 1, 512, 4, [(12, 10), (6, 5), (3, 3)], 1),      # 64 channels per head, non-square levels
}
for seed, (name, (C, M, NL, F_, P, shapes, N)) in enumerate(CASES.items()):
    g = torch.Generator().manual_seed(900 + seed)
    L = len(shapes)
    torch.manual_seed(77 + seed)                                                   # xavier_uniform_ / normal_ draw from the global generator
    mod = RefEncoderOnly(d_model=C, nhead=M, num_encoder_layers=NL, dim_feedforward=F_, dropout=0.1, activation="relu",
                         num_feature_levels=L, enc_n_points=P).double().eval()
    with torch.no_grad():
        for k, p in mod.named_parameters():
            if k.endswith("sampling_offsets.bias"):
                p.add_((torch.randn(p.shape, generator=g) * 0.3).double())         # the reference's ring initialisation, perturbed
            elif "norm" in k:
                p.add_((torch.randn(p.shape, generator=g) * 0.2).double())
            elif k.endswith("sampling_offsets.weight"):
                p.copy_((torch.randn(p.shape, generator=g) * 0.3).double())
            elif k.endswith("attention_weights.weight") or k.endswith("bias"):
                p.copy_((torch.randn(p.shape, generator=g) * 0.2).double())
            # the remaining weights keep the reference's xavier / normal initialisation (drawn from torch's global generator)
    sd = {k: v.detach().float() for k, v in mod.state_dict().items()}
    mod.load_state_dict({k: v.double() for k, v in sd.items()})                    # parameters are exactly fp32-representable
    srcs = [torch.randn(N, C, h, w, generator=g) for h, w in shapes]
    poss = [torch.randn(N, C, h, w, generator=g) for h, w in shapes]
    with torch.no_grad():
        memory, spatial_shapes, level_start_index = mod([s.double() for s in srcs], [p.double() for p in poss])
    blob = {"memory": memory.float().numpy(), "shapes": spatial_shapes.numpy(), "starts": level_start_index.numpy(),
            "meta": np.array([C, M, NL, F_, P, L], dtype=np.int64)}
    for i in range(L):
        blob[f"src{i}"] = srcs[i].numpy()
        blob[f"pos{i}"] = poss[i].numpy()
    for k, v in sd.items():
        blob["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"deform_encoder_{name}.npz"), **blob)
    print(name, tuple(memory.shape), float(memory.abs().mean()), len(sd))
