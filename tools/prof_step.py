#!/usr/bin/env python
"""Minimal driver for ncu captures: N forwards of one chunk of the chosen encoder (no e2e leg, no CPU baseline).

    python tools/prof_step.py [--model vit_h] [--batch 8] [--steps 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import iuvl_b200 as ib  # noqa: E402
from iuvl_b200.encoder import build_encoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="vit_h")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()
cfg = ib.PRESETS[args.model]
enc = build_encoder(cfg)
enc.load_state_dict(ib.make_state_dict(cfg, 1234))
enc.to("cuda:0")
enc.out_dtype = torch.bfloat16
enc.max_chunk = args.batch
x = torch.randn(args.batch, 3, cfg.img_size, cfg.img_size, device="cuda:0")
with torch.no_grad():
    for _ in range(args.steps):
        out = enc(x)
torch.cuda.synchronize()
print("ok", {k: tuple(v.shape) for k, v in out.items()})
