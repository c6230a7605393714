#!/usr/bin/env python
"""Timing of the pixel decoder at the step1.yaml geometry (transformer_encoder_deform.py:289-311): conv_dim = mask_dim = 512, 8 heads,
6 encoder layers, d_ffn 1024, fed with the four maps the ViT encoder produces for 1024^2 images (res2 128x256^2 ... res5 1024x32^2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200 import cabi  # noqa: E402
from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = "cuda"
mod = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=6,
                               conv_dim=512, mask_dim=512, norm="GN").to(dev).eval()
lib = cabi.lib()
with torch.no_grad():
    for layer in mod.transformer.encoder.layers:
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    feats = {f"res{2 + i}": torch.randn(N, c, 256 >> i, 256 >> i, device=dev).bfloat16() for i, c in enumerate((128, 256, 512, 1024))}
    for precision in ("bf16",):
        mod.precision = precision
        for _ in range(2):
            mask, multi = mod(feats)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        n0 = lib.svb_launch_count()
        e0.record()
        for _ in range(3):
            mask, multi = mod(feats)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        S = 32 * 32 + 64 * 64 + 128 * 128
        fl = 2.0 * N * (32 * 32 * 1024 * 512 + 64 * 64 * 512 * 512 + 128 * 128 * 256 * 512 + 65536 * (128 * 512 + 9 * 512 * 512 + 512 * 512)
                        + 6 * S * 512 * (512 + 288 + 512 + 1024 + 1024))
        print(f"{precision}: {ms:8.2f} ms per forward for {N} images ({N / ms * 1e3:.0f} images/s), {(lib.svb_launch_count() - n0) // 3} launches; "
              f"GEMM work {fl / 1e12:.2f} TFLOP ({fl / ms / 1e9:.0f} TF/s over the whole decoder); mask_features {tuple(mask.shape)}")
