#!/usr/bin/env python
"""Timing of the multi-scale deformable attention forward at the step1.yaml geometry: 1024^2 input -> levels 128^2, 64^2, 32^2
(res3..res5), 8 heads x 64 channels, 4 points, queries = all 21504 positions.  Reports the gather bandwidth (algorithmic bytes:
4 taps x 16 B x channel groups per sample + locations/weights + output) against the measured HBM copy peak for orientation (the
value maps are L2-resident)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200.msda import ms_deform_attn_forward  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shapes = [(128, 128), (64, 64), (32, 32)]
M, D, P, L = 8, 64, 4, 3
S = sum(h * w for h, w in shapes)
st = [0, 128 * 128, 128 * 128 + 64 * 64]
dev = "cuda"
for dtype in (torch.bfloat16, torch.float32):
    value = torch.randn(N, S, M, D, device=dev).to(dtype)
    ref_pts = torch.rand(N, S, 1, 1, 1, 2, device=dev)
    loc = (ref_pts + 0.05 * torch.randn(N, S, M, L, P, 2, device=dev)).contiguous()       # offsets around a reference point
    aw = torch.softmax(torch.randn(N, S, M, L * P, device=dev), -1).reshape(N, S, M, L, P).contiguous()
    sh, stt = torch.tensor(shapes), torch.tensor(st)
    with torch.no_grad():
        for _ in range(3):
            out = ms_deform_attn_forward(value, sh, stt, loc, aw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            out = ms_deform_attn_forward(value, sh, stt, loc, aw)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    es = value.element_size()
    gather = N * S * M * L * P * 4 * D * es
    other = N * S * M * L * P * 12 + N * S * M * D * es
    print(f"{str(dtype):16s} N={N}: {ms * 1e3:8.1f} us   gather {gather / ms / 1e6:8.1f} GB/s   (+ loc/weights/out {other / 1e6:.0f} MB)"
          f"   compulsory HBM {(value.numel() * es + other) / ms / 1e6:7.1f} GB/s")
