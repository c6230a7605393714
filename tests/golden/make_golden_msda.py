"""Golden fixtures of the multi-scale deformable attention forward, FROM THE REFERENCE'S OWN FUNCTION.

Run in the build container only:  python tests/golden/make_golden_msda.py

``/root/reference/modeling/vision/encoder/ops/functions/ms_deform_attn_func.py`` cannot be imported (it requires the compiled
MultiScaleDeformableAttention extension at import time, :21-29), so the SOURCE TEXT of its pure-PyTorch implementation
``ms_deform_attn_core_pytorch`` (:52-72) — the implementation the reference's own ``ops/test.py`` checks its CUDA kernel
against — is extracted and executed unmodified.  Inputs (seeded) and fp64 outputs are stored as tests/golden/msda_<case>.npz.
"""
import os
import re

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/modeling/vision/encoder/ops/functions/ms_deform_attn_func.py"

text = open(SRC).read()
m = re.search(r"^def ms_deform_attn_core_pytorch\(.*?(?=^\S|\Z)", text, re.S | re.M)
ns = {"torch": torch, "F": F}
exec(m.group(0), ns)                                   # the reference function, unmodified
ref_fn = ns["ms_deform_attn_core_pytorch"]

# name -> (N, M, D, Lq, P, shapes, location range)
CASES = {
    "toy": (1, 2, 2, 2, 2, [(6, 4), (3, 2)], (0.0, 1.0)),                 # the geometry of ops/test.py:24-29
    "small": (2, 4, 8, 37, 4, [(12, 9), (6, 5), (3, 3)], (-0.2, 1.2)),      # out-of-range locations: zero padding, edge taps
    "heads8": (1, 8, 64, 50, 4, [(8, 8), (4, 4), (2, 2)], (-0.05, 1.05)),    # the step1.yaml head geometry, tiny maps
}
for seed, (name, (N, M, D, Lq, P, shapes, (lo, hi))) in enumerate(CASES.items()):
    g = torch.Generator().manual_seed(100 + seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    # inputs are fp32-representable (stored as fp32), the reference function runs on their fp64 copies
    value = torch.randn(N, S, M, D, generator=g, dtype=torch.float32)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float32) * (hi - lo) + lo
    aw = torch.rand(N, Lq, M, L, P, generator=g, dtype=torch.float32) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)               # as ops/test.py:37-38
    out = ref_fn(value.double(), torch.as_tensor(shapes), loc.double(), aw.double())
    np.savez_compressed(os.path.join(HERE, f"msda_{name}.npz"), value=value.numpy(), loc=loc.numpy(), aw=aw.numpy(),
                        shapes=np.array(shapes, dtype=np.int64), out=out.numpy())
    print(name, tuple(out.shape), float(out.abs().mean()))
