"""Scope row N3: forward time of the encoder on a canvas other than 1024 x 1024 (default 1024 x 2048, the reference's COCO evaluation pad)
next to the 1024 x 1024 time; SVB_ATTN_EXT=0 selects the round-1 path (fp32-math attention kernel for such grids)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iuvl_b200 as ib
from iuvl_b200.encoder import build_encoder
model = sys.argv[1] if len(sys.argv) > 1 else "vit_b"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1024, 2048)
B = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfg = ib.PRESETS[model]
enc = build_encoder(cfg); enc.load_state_dict(ib.make_state_dict(cfg, 1234)); enc.to("cuda"); enc.out_dtype = torch.bfloat16
for (h, w) in ((1024, 1024), (H, W)):
    x = torch.randn(B, 3, h, w, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            enc(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            enc(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{model} {h}x{w} batch {B} (SVB_ATTN_EXT={os.environ.get('SVB_ATTN_EXT', '1')}): {ms:.2f} ms per forward, {B / ms * 1e3:.1f} images/s, "
          f"{B * h * w / 1048576 / ms * 1e3:.1f} Mpixel-images/s")
