#!/usr/bin/env python
"""Timing of the whole MSDeformAttn module forward (4 GEMMs + fused sampling kernel) at the step1.yaml geometry:
d_model 512, 8 heads, 3 levels (128^2, 64^2, 32^2), 4 points, queries = all 21504 positions, N images."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200.msda import MSDeformAttn  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shapes = [(128, 128), (64, 64), (32, 32)]
S = sum(h * w for h, w in shapes)
st = [0, 128 * 128, 128 * 128 + 64 * 64]
dev = "cuda"
mod = MSDeformAttn(512, 3, 8, 4).to(dev)
with torch.no_grad():
    mod.sampling_offsets.weight.normal_(0, 0.05)
    mod.attention_weights.weight.normal_(0, 0.05)
    x = torch.randn(N, S, 512, device=dev)
    ref = torch.rand(N, S, 3, 2, device=dev)
    sh, stt = torch.tensor(shapes), torch.tensor(st)
    for precision in ("bf16", "fp32"):
        mod.precision = precision
        xin = x.bfloat16() if precision == "bf16" else x
        for _ in range(2):
            out = mod(xin, ref, xin, sh, stt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            out = mod(xin, ref, xin, sh, stt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * N * S * 512 * (512 + 288 + 512)
        print(f"{precision}: {ms:8.3f} ms per layer for {N} images ({N / ms * 1e3:.0f} images/s per layer); linears {fl / 1e9:.1f} GFLOP")
