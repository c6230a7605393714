#!/usr/bin/env python
"""BASELINE config 5 shape on one GPU: ViT-H encoder -> MSDeformAttn pixel decoder -> X-Decoder mask path, N synthetic 1024^2 images,
random-init weights, bf16 hand-overs.  Prints the time of each stage and the pipeline's images/s (device-resident inputs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import iuvl_b200 as ib  # noqa: E402
from iuvl_b200.encoder import build_encoder  # noqa: E402
from iuvl_b200.mask_head import XDecoderMaskPath  # noqa: E402
from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = "cuda"
cfg = ib.PRESETS["vit_h"]
enc = build_encoder(cfg)
enc.load_state_dict(ib.make_state_dict(cfg, 1234))
enc.to(dev)
enc.out_dtype = torch.bfloat16
dec = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=6,
                               conv_dim=512, mask_dim=512, norm="GN").to(dev).eval()
path = XDecoderMaskPath(512, 512, 101, 8, 2048).to(dev).eval()
with torch.no_grad():
    for layer in dec.transformer.encoder.layers:
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    x = torch.randn(N, 3, 1024, 1024, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for rep in range(3):
        ev[0].record()
        feats = enc(x)
        ev[1].record()
        _, multi, extra = dec(feats, rows_out=True)
        ev[2].record()
        out = path(multi, None, mask_rows=extra["mask_rows"], mask_shape=extra["mask_shape"])
        ev[3].record()
        torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    print(f"{N} images: encoder {t[0]:.1f} ms, pixel decoder {t[1]:.1f} ms, mask path {t[2]:.1f} ms -> {N / sum(t) * 1e3:.1f} images/s; "
          f"pred_masks {tuple(out['pred_masks'].shape)}")
