"""Golden fixtures of the MSDeformAttn pixel decoder, FROM THE UNMODIFIED REFERENCE CLASSES.

Run in the build container only:  python tests/golden/make_golden_pixel_decoder.py

``transformer_encoder_deform.py`` imports detectron2 / fvcore at module level (absent here).  As in make_golden_deform_encoder.py the
SOURCE TEXT of its classes (``MSDeformAttnPixelDecoder`` :164-359 and the three encoder classes :23-161) is cut out with ``ast`` and
executed unmodified in a namespace that holds what they use: torch, the reference's own ``MSDeformAttn`` (pure-PyTorch branch),
``_get_clones`` / ``_get_activation_fn``, ``PositionEmbeddingSine`` (``modeling/modules/position_encoding.py``, importable);
``@configurable`` (``modeling/utils/config.py``: needs omegaconf only to recognise a cfg object) is the identity for the keyword
construction used here — plus RESTATEMENTS of the three detectron2 / fvcore names the class
touches, by their published definitions (detectron2 v0.6 ``layers/wrappers.py::Conv2d``: ``F.conv2d`` -> ``norm`` -> ``activation``;
``layers/batch_norm.py::get_norm("GN", C)`` = ``nn.GroupNorm(32, C)``; fvcore ``c2_xavier_fill`` = ``kaiming_uniform_(a=1)`` + zero bias).
That third-party part is "parity unpinned" beyond those definitions.  The module runs in fp32 / eval() (its forward casts the features with `.float()`, :320,345).
"""
import ast
import importlib.util
import os
import sys
import types
import typing

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

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
posenc = _load("ref_position_encoding", "/root/reference/modeling/modules/position_encoding.py")


class Conv2d(nn.Conv2d):                       # detectron2.layers.Conv2d (restated)
    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def get_norm(norm, out_channels):              # detectron2.layers.get_norm (restated for the values the class passes)
    if norm is None or norm == "":
        return None
    assert norm == "GN"
    return nn.GroupNorm(32, out_channels)


weight_init = types.SimpleNamespace()


def _c2_xavier_fill(module):                   # fvcore.nn.weight_init.c2_xavier_fill (restated)
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


weight_init.c2_xavier_fill = _c2_xavier_fill


class _NoAutocast:                              # `@autocast(enabled=False)` of the forward: a no-op on the CPU
    def __init__(self, enabled=True):
        pass

    def __call__(self, fn):
        return fn


text = open(ENC + "/transformer_encoder_deform.py").read()
wanted = ("MSDeformAttnTransformerEncoderOnly", "MSDeformAttnTransformerEncoderLayer", "MSDeformAttnTransformerEncoder", "MSDeformAttnPixelDecoder")
ns = {"torch": torch, "nn": nn, "F": F, "np": np, "normal_": nn.init.normal_, "MSDeformAttn": RefMSDeformAttn, "_get_clones": blocks._get_clones,
      "_get_activation_fn": blocks._get_activation_fn, "Conv2d": Conv2d, "get_norm": get_norm, "weight_init": weight_init,
      "PositionEmbeddingSine": posenc.PositionEmbeddingSine, "configurable": (lambda init: init), "autocast": _NoAutocast,
      "ShapeSpec": object, "Dict": typing.Dict, "List": typing.List, "Optional": typing.Optional, "Union": typing.Union,
      "Callable": typing.Callable, "Tuple": typing.Tuple}
for node in ast.parse(text).body:
    if isinstance(node, ast.ClassDef) and node.name in wanted:
        exec(compile(ast.Module(body=[node], type_ignores=[]), "transformer_encoder_deform.py", "exec"), ns)
RefPixelDecoder = ns["MSDeformAttnPixelDecoder"]

# name -> (conv_dim, mask_dim, heads, layers, d_ffn, batch, res2 side)   (res3/4/5 sides = res2 side / 2, 4, 8)
CASES = {
    "small": (64, 32, 4, 2, 128, 2, 32),
    "wide": (128, 64, 2, 1, 256, 1, 48),          # 64 channels per head, non-power-of-two maps (48, 24, 12, 6)
}
for seed, (name, (C, MD, M, NL, F_, N, side)) in enumerate(CASES.items()):
    g = torch.Generator().manual_seed(1300 + seed)
    torch.manual_seed(91 + seed)
    mod = RefPixelDecoder(input_shape=None, transformer_dropout=0.0, transformer_nheads=M, transformer_dim_feedforward=F_,
                          transformer_enc_layers=NL, conv_dim=C, mask_dim=MD, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                          common_stride=4).eval()
    with torch.no_grad():
        for k, p in mod.named_parameters():
            if k.endswith("sampling_offsets.bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.3)
            elif "norm" in k or k.startswith("input_proj") and k.split(".")[2] == "1":
                p.add_(torch.randn(p.shape, generator=g) * 0.2)          # every GroupNorm / LayerNorm away from (1, 0)
            elif k.endswith("sampling_offsets.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif k.endswith("attention_weights.weight") or k.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)
    sd = {k: v.detach().float() for k, v in mod.state_dict().items()}
    gf = torch.Generator().manual_seed(1400 + seed)      # the features are NOT stored: tests regenerate them from this seed
    feats = {f"res{2 + i}": torch.randn(N, c, side >> i, side >> i, generator=gf) for i, c in enumerate((128, 256, 512, 1024))}
    with torch.no_grad():
        mask, multi = mod(dict(feats))
    blob = {"mask_features": mask.float().numpy(), "meta": np.array([C, MD, M, NL, F_], dtype=np.int64)}
    for i, m in enumerate(multi):
        blob[f"multi{i}"] = m.float().numpy()
    blob["feat_seed"] = np.array([1400 + seed, N, side], dtype=np.int64)
    for k, v in sd.items():
        blob["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"pixel_decoder_{name}.npz"), **blob)
    print(name, tuple(mask.shape), [tuple(m.shape) for m in multi], float(mask.abs().mean()), len(sd))


# ---- the step1.yaml geometry (conv_dim = mask_dim = 512, 8 heads, d_ffn 1024, 6 encoder layers; configs/step1.yaml) on small maps.  The
# 11.5 M weights are not stored: tests.util.seeded_state_dict regenerates them; the outputs are stored as 32768 samples each. ----
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.util import sampled, seeded_state_dict  # noqa: E402
C, MD, M, NL, F_, N, side, SEED = 512, 512, 8, 6, 1024, 1, 32, 4100
mod = RefPixelDecoder(input_shape=None, transformer_dropout=0.0, transformer_nheads=M, transformer_dim_feedforward=F_,
                      transformer_enc_layers=NL, conv_dim=C, mask_dim=MD, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                      common_stride=4).eval()
mod.load_state_dict(seeded_state_dict(list(mod.state_dict().items()), SEED), strict=True)
gf = torch.Generator().manual_seed(SEED + 1)
feats = {f"res{2 + i}": torch.randn(N, c, side >> i, side >> i, generator=gf) for i, c in enumerate((128, 256, 512, 1024))}
with torch.no_grad():
    mask, multi = mod(dict(feats))
blob = {"meta": np.array([C, MD, M, NL, F_, N, side, SEED], dtype=np.int64)}
for name, t in [("mask_features", mask)] + [(f"multi{i}", m) for i, m in enumerate(multi)]:
    idx, val = sampled(t, 32768, 7)
    blob[name + ".idx"], blob[name + ".val"], blob[name + ".shape"] = idx.numpy(), val.numpy(), np.array(t.shape, dtype=np.int64)
np.savez_compressed(os.path.join(HERE, "pixel_decoder_step1.npz"), **blob)
print("step1", tuple(mask.shape), [tuple(m.shape) for m in multi], float(mask.abs().mean()))
