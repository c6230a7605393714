#!/usr/bin/env python
"""Print the rel-L2 of every tap / output of one golden case (default vit_h_std, bf16): the margins under the 1e-2 bar."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests.test_gpu_encoder import _setup  # noqa: E402
from tests.util import sampled_rel_l2  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "vit_h_std"
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
g, cfg, sd, x, enc = _setup(case)
enc.precision = precision
enc.enable_taps(True)
with torch.no_grad():
    out = enc(x.to("cuda"))
torch.cuda.synchronize()
names = ["embed"] + [f"block{i}" for i in range(cfg.depth)]
taps = [sampled_rel_l2(enc.read_tap(i - 1), g, "tap." + n) for i, n in enumerate(names)]
print(case, precision, "taps:", " ".join(f"{e:.1e}" for e in taps))
print(case, precision, "outs:", {k: f"{sampled_rel_l2(out[k].float(), g, 'out.' + k):.2e}" for k in ("res2", "res3", "res4", "res5")})
