#!/usr/bin/env python
"""Small end-to-end run for `compute-sanitizer --tool memcheck`: tiny80 encoder (fp32 and bf16 paths, square and 512 x 1024
canvases, uint8 entry) + the fused GEMM modes + the deformable-attention kernel, on shapes that exercise every edge path."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import iuvl_b200 as ib  # noqa: E402
from iuvl_b200.encoder import build_encoder  # noqa: E402
from iuvl_b200.msda import ms_deform_attn_forward  # noqa: E402

dev = "cuda"
cfg = ib.PRESETS["tiny80"]
enc = build_encoder(cfg)
enc.load_state_dict(ib.make_state_dict(cfg, 1, rel_std=0.1))
enc.to(dev)
with torch.no_grad():
    for precision in ("bf16", "fp32"):
        enc.precision = precision
        out = enc(ib.make_images(2, cfg, 0).to(dev))
        out = enc(ib.make_images(1, cfg, 0, hw=(512, 1024)).to(dev))
        imgs = [torch.randint(0, 256, (3, 700, 999), dtype=torch.uint8, device=dev), torch.randint(0, 256, (3, 5, 1024), dtype=torch.uint8, device=dev)]
        out = enc.forward_uint8(imgs, [123.675, 116.28, 103.53], [58.395, 57.12, 57.375])
    shapes = [(9, 7), (4, 5)]
    v = torch.randn(2, 83, 3, 16, device=dev)
    loc = torch.rand(2, 31, 3, 2, 4, 2, device=dev) * 1.4 - 0.2
    aw = torch.rand(2, 31, 3, 2, 4, device=dev)
    o = ms_deform_attn_forward(v, torch.tensor(shapes), torch.tensor([0, 63]), loc, aw)
    o = ms_deform_attn_forward(v.bfloat16(), torch.tensor(shapes), torch.tensor([0, 63]), loc, aw)
torch.cuda.synchronize()
print("sanitize run ok", float(out["res5"].float().abs().mean()), float(o.float().abs().mean()))
