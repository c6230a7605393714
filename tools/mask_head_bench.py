#!/usr/bin/env python
"""Timing of the mask branch of the prediction heads at the step1.yaml geometry: 101 queries, hidden = mask_dim = 512, 8 heads,
mask_features 512 x 256^2 per image, attention-mask sizes 32^2 / 64^2 / 128^2 (one call per decoder layer, xdecoder.py:296-329)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200.mask_head import MaskPredictionHead  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = "cuda"
head = MaskPredictionHead(512, 512, 101, 8).to(dev).eval()
with torch.no_grad():
    out = torch.randn(101, N, 512, device=dev)
    mf = torch.randn(N, 512, 256, 256, device=dev).bfloat16()
    for tgt in ((32, 32), (64, 64), (128, 128)):
        for _ in range(2):
            res = head(out, mf, tgt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            res = head(out, mf, tgt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * N * 101 * 512 * 65536
        print(f"target {tgt}: {ms:7.3f} ms per call for {N} images; mask-logit GEMMs {fl / 1e9:.0f} GFLOP; outputs_mask {tuple(res['outputs_mask'].shape)}")

# the masked cross-attention layer of a decoder layer (interface/modules.py:95-106) at the three memory sizes
from iuvl_b200.mask_head import CrossAttentionLayer  # noqa: E402
layer = CrossAttentionLayer(512, 8).to(dev).eval()
with torch.no_grad():
    tgt, qpos = torch.randn(101, N, 512, device=dev), torch.randn(101, N, 512, device=dev)
    for side in (32, 64, 128):
        hw = side * side
        mem, pos = torch.randn(hw, N, 512, device=dev), torch.randn(hw, N, 512, device=dev)
        mask = torch.rand(N * 8, 101, hw, device=dev) < 0.5
        for _ in range(2):
            layer(tgt, mem, memory_mask=mask, pos=pos, query_pos=qpos)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            layer(tgt, mem, memory_mask=mask, pos=pos, query_pos=qpos)
        e1.record()
        torch.cuda.synchronize()
        print(f"cross-attention layer, memory {side}^2: {e0.elapsed_time(e1) / 5:7.3f} ms per call for {N} images")

# the whole mask path of the X-Decoder predictor (9 layers) on the pixel decoder's outputs
from iuvl_b200.mask_head import XDecoderMaskPath  # noqa: E402
path = XDecoderMaskPath(512, 512, 101, 8, 2048).to(dev).eval()
with torch.no_grad():
    xs = [torch.randn(N, 512, s, s, device=dev) for s in (32, 64, 128)]
    for _ in range(2):
        res = path(xs, mf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        res = path(xs, mf)
    e1.record()
    torch.cuda.synchronize()
    print(f"X-Decoder mask path (9 layers, 10 prediction-head calls): {e0.elapsed_time(e1) / 3:7.2f} ms per forward for {N} images; pred_masks {tuple(res['pred_masks'].shape)}")
