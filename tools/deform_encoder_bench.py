#!/usr/bin/env python
"""Timing of the deformable-attention encoder (MSDeformAttnTransformerEncoderOnly) at the step1.yaml geometry
(transformer_encoder_deform.py:289-311): d_model 512, 8 heads, 6 layers, d_ffn 1024, levels 32^2, 64^2, 128^2 (res5 -> res3), 4 points."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200.msda import MSDeformAttnTransformerEncoderOnly  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shapes = [(32, 32), (64, 64), (128, 128)]
dev = "cuda"
mod = MSDeformAttnTransformerEncoderOnly(512, 8, 6, 1024, 0.0, "relu", 3, 4).to(dev).eval()
with torch.no_grad():
    for layer in mod.encoder.layers:
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    srcs = [torch.randn(N, 512, h, w, device=dev) for h, w in shapes]
    poss = [torch.randn(N, 512, h, w, device=dev) for h, w in shapes]
    S = sum(h * w for h, w in shapes)
    for precision in ("bf16", "fp32"):
        mod.precision = precision
        for _ in range(2):
            out = mod(srcs, poss)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            out = mod(srcs, poss)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 6 * 2.0 * N * S * 512 * (512 + 288 + 512 + 1024 + 1024)
        print(f"{precision}: {ms:8.3f} ms per 6-layer encoder for {N} images ({N / ms * 1e3:.0f} images/s); linears {fl / 1e9:.1f} GFLOP "
              f"({fl / ms / 1e9:.0f} TF/s over the whole encoder)")
