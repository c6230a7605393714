#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM through the C ABI (svb_linear) on the encoder's shapes, back to back (sustained
clocks), with torch.matmul (cuBLAS) on the same shapes as a yardstick only.

    python tools/gemm_bench.py [--images 8] [--reps 30] [--model vit_h] [--check]

Environment knobs are read once by the library (SVB_GEMM_CLUSTER=2|4, SVB_GEMM_IMPL): run one process per variant.
"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import iuvl_b200 as ib  # noqa: E402
from iuvl_b200 import cabi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=8)
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--model", default="vit_h")
ap.add_argument("--check", action="store_true")
ap.add_argument("--no-cublas", action="store_true")
args = ap.parse_args()
cfg = ib.PRESETS[args.model]
D, T = cfg.embed_dim, 4096
M = args.images * T
dev = "cuda"
lib = cabi.lib()


def linear(A, W, bias, gelu, resid, out):
    Mx, K = A.shape
    N = W.shape[0]
    rc = lib.svb_linear(cabi.MODE_BF16, A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), Mx, N, K, cabi.ptr(bias), int(gelu),
                        cabi.ptr(resid), resid.stride(0) if resid is not None else 0, 0, out.data_ptr(),
                        cabi.DTYPE_BF16 if out.dtype == torch.bfloat16 else cabi.DTYPE_F32, out.stride(0), None, 0, 0, 0,
                        cabi.stream_ptr())
    cabi.check(rc, "svb_linear")


shapes = [("qkv", M, 3 * D, D, False, False, torch.bfloat16), ("proj", M, D, D, False, False, torch.bfloat16),
          ("lin1", M, 4 * D, D, True, False, torch.bfloat16), ("lin2", M, D, 4 * D, False, True, torch.float32)]
tot_ms, tot_fl = 0.0, 0.0
for name, m, n, k, gelu, resid, odt in shapes:
    g = torch.Generator(device="cpu").manual_seed(n + k)
    A = (torch.randn(m, k, device=dev)).bfloat16()
    W = (torch.randn(n, k, device=dev) / math.sqrt(k)).bfloat16()
    bias = torch.randn(n, device=dev)
    out = torch.zeros(m, n, dtype=odt, device=dev)
    R = out if resid else None
    if args.check:
        out.zero_()
        linear(A, W, bias, gelu, R, out)
        torch.cuda.synchronize()
        idx = torch.randint(0, m, (2048,), device=dev)
        ref = A[idx].double() @ W.double().t() + bias.double()
        if gelu:
            ref = 0.5 * ref * (1 + torch.erf(ref / math.sqrt(2)))
        err = ib.rel_l2(out[idx], ref)
        print(f"  check {name}: rel_l2 {err:.2e}")
        assert err < 5e-3, err
    for _ in range(3):
        linear(A, W, bias, gelu, R, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.reps):
        linear(A, W, bias, gelu, R, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    fl = 2.0 * m * n * k
    line = f"{name:5s} M={m} N={n} K={k}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TF/s"
    if not args.no_cublas:
        Wt = W.t().contiguous()
        o2 = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            torch.matmul(A, Wt, out=o2)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.reps):
            torch.matmul(A, Wt, out=o2)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / args.reps
        line += f"   | cuBLAS (no epilogue) {ms2 * 1e3:8.1f} us {fl / ms2 / 1e9:7.1f} TF/s"
    print(line, flush=True)
    tot_ms += ms
    tot_fl += fl
print(f"block linears: {tot_ms * 1e3:.1f} us per block of {args.images} images, {tot_fl / tot_ms / 1e9:.1f} TF/s "
      f"(SVB_GEMM_CLUSTER={os.environ.get('SVB_GEMM_CLUSTER', 'default')})")
