#!/usr/bin/env python
"""Minimal driver for ncu captures of the rows N1 / N2 / N4 kernels at the step1.yaml geometry: uint8 staging, pixel decoder, mask path
(N images of 1024 x 1024: maps of 256^2 .. 32^2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C  # noqa: E402
import torch  # noqa: E402

from iuvl_b200 import cabi  # noqa: E402
from iuvl_b200.mask_head import XDecoderMaskPath  # noqa: E402
from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = "cuda"
torch.manual_seed(0)
dec = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=6,
                               conv_dim=512, mask_dim=512, norm="GN").to(dev).eval()
path = XDecoderMaskPath(512, 512, 101, 8, 2048).to(dev).eval()
with torch.no_grad():
    for layer in dec.transformer.encoder.layers:
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    feats = {f"res{2 + i}": torch.randn(N, c, 256 >> i, 256 >> i, device=dev).bfloat16() for i, c in enumerate((128, 256, 512, 1024))}
    for _ in range(2):
        _, multi, extra = dec(feats, rows_out=True)
        out = path(multi, None, mask_rows=extra["mask_rows"], mask_shape=extra["mask_shape"])
    imgs = [torch.randint(0, 256, (3, 1024, 1024), dtype=torch.uint8, device=dev) for _ in range(16)]
    dst = torch.empty(16 * 4096, 768, dtype=torch.bfloat16, device=dev)
    ptrs = (C.c_void_p * 16)(*[t.data_ptr() for t in imgs])
    hs, wsz = (C.c_int * 16)(*([1024] * 16)), (C.c_int * 16)(*([1024] * 16))
    mean, std = (C.c_float * 3)(123.675, 116.28, 103.53), (C.c_float * 3)(58.395, 57.12, 57.375)
    for _ in range(2):
        cabi.check(cabi.lib().svb_stage_images_u8(ptrs, hs, wsz, 16, 3, 1024, 16, mean, std, dst.data_ptr(), cabi.DTYPE_BF16, cabi.stream_ptr()))
torch.cuda.synchronize()
print("ok", tuple(out["pred_masks"].shape))
