"""Golden fixtures of the whole MSDeformAttn module forward, FROM THE UNMODIFIED REFERENCE MODULE.

Run in the build container only:  python tests/golden/make_golden_msda_module.py

The reference package needs the compiled ``MultiScaleDeformableAttention`` extension at import time
(``ops/functions/ms_deform_attn_func.py:21-29``).  A stub module of that name (whose functions raise) lets the package import;
``MSDeformAttn.forward`` then takes its own ``except:`` branch, the pure-PyTorch core (``ops/modules/ms_deform_attn.py:117-122``).
Parameters are randomised (the reference initialises the query Linears' weights to zero, which would hide them), inputs seeded;
everything is stored in fp32, the module runs in fp64.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OPS = "/root/reference/modeling/vision/encoder/ops"

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

# name -> (d_model, heads, points, shapes, N, Lq, ref_dim, padding mask?)
CASES = {
    "points": (64, 4, 4, [(6, 5), (3, 3)], 2, 19, 2, False),
    "boxes_masked": (128, 8, 2, [(8, 8), (4, 4), (2, 2)], 1, 33, 4, True),
    "heads64": (256, 4, 4, [(8, 8), (4, 4), (2, 2)], 1, 84, 2, False),         # 64 channels per head as in step1.yaml (512 = 8 x 64)
}
for seed, (name, (C, M, P, shapes, N, Lq, rd, masked)) in enumerate(CASES.items()):
    g = torch.Generator().manual_seed(500 + seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    mod = RefMSDeformAttn(d_model=C, n_levels=L, n_heads=M, n_points=P).double()
    with torch.no_grad():
        for k, p in mod.named_parameters():
            scale = 0.5 if "sampling_offsets.weight" in k else (1.0 / np.sqrt(C) if k.endswith("weight") else 0.3)
            if k == "sampling_offsets.bias":
                p.add_((torch.randn(p.shape, generator=g) * 0.3).double())         # keep the reference's ring initialisation, perturbed
            else:
                p.copy_((torch.randn(p.shape, generator=g) * scale).double())
    sd = {k: v.detach().float() for k, v in mod.state_dict().items()}
    mod.load_state_dict({k: v.double() for k, v in sd.items()})                    # parameters are exactly fp32-representable
    query = torch.randn(N, Lq, C, generator=g)
    inp = torch.randn(N, S, C, generator=g)
    if rd == 2:
        ref = torch.rand(N, Lq, L, 2, generator=g)
    else:
        ref = torch.cat([torch.rand(N, Lq, L, 2, generator=g), torch.rand(N, Lq, L, 2, generator=g) * 0.5 + 0.05], -1)
    mask = (torch.rand(N, S, generator=g) < 0.15) if masked else None
    starts = [0]
    for h, w in shapes[:-1]:
        starts.append(starts[-1] + h * w)
    with torch.no_grad():
        out = mod(query.double(), ref.double(), inp.double(), torch.tensor(shapes), torch.tensor(starts), mask)
    blob = {"query": query.numpy(), "inp": inp.numpy(), "ref": ref.numpy(), "shapes": np.array(shapes, dtype=np.int64),
            "out": out.numpy(), "meta": np.array([C, M, P, L], dtype=np.int64), "mask": (mask.numpy() if masked else np.zeros(0, dtype=bool))}
    for k, v in sd.items():
        blob["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"msda_module_{name}.npz"), **blob)
    print(name, tuple(out.shape), float(out.abs().mean()))
