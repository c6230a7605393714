#!/usr/bin/env python
"""Timing of the GEMM epilogue variants of the LayerNorm-folded bf16 path (svb_linear_fused), ViT-H shapes, 8 images."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200 import cabi  # noqa: E402

dev = "cuda"
lib = cabi.lib()
D, M = 1280, 8 * 4096
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
parts = (D + 127) // 128


def run(name, A, W, bias, out, gelu=False, resid=None, ln=None, out2=None, stat=None):
    Mx, K = A.shape
    N = W.shape[0]
    ln_stats, ln_c = ln if ln is not None else (None, None)

    def call():
        rc = lib.svb_linear_fused(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), Mx, N, K, cabi.ptr(bias), int(gelu), cabi.ptr(resid),
                                  resid.stride(0) if resid is not None else 0, 0, out.data_ptr(),
                                  cabi.DTYPE_BF16 if out.dtype == torch.bfloat16 else cabi.DTYPE_F32, out.stride(0), cabi.ptr(ln_stats),
                                  cabi.ptr(ln_c), D if ln is not None else 0, 1e-6, cabi.ptr(out2), D if out2 is not None else 0,
                                  cabi.ptr(stat), 0, 0, cabi.stream_ptr())
        cabi.check(rc, name)
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} {ms * 1e3:8.1f} us  {2.0 * Mx * N * K / ms / 1e9:7.1f} TF/s", flush=True)


def mk(m, n, k):
    A = torch.randn(m, k, device=dev).bfloat16()
    W = (torch.randn(n, k, device=dev) / math.sqrt(k)).bfloat16()
    return A, W, torch.randn(n, device=dev)


X = torch.randn(M, D, device=dev)
Xb = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
st = torch.zeros(M, parts, 2, device=dev)
st[:, :, 1] = 128.0
for nm, k in (("proj", D), ("lin2", 4 * D)):
    A, W, b = mk(M, D, k)
    ob = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    run(f"{nm} bf16 out", A, W, b, ob)
    run(f"{nm} fp32 resid in place", A, W, b, X, resid=X)
    run(f"{nm} + bf16 copy", A, W, b, X, resid=X, out2=Xb)
    run(f"{nm} + row stats", A, W, b, X, resid=X, stat=st)
    run(f"{nm} + bf16 copy + row stats", A, W, b, X, resid=X, out2=Xb, stat=st)
for nm, n, gelu in (("qkv", 3 * D, False), ("lin1", 4 * D, True)):
    A, W, b = mk(M, n, D)
    c = torch.randn(n, device=dev)
    o = torch.empty(M, n, dtype=torch.bfloat16, device=dev)
    run(f"{nm} plain", A, W, b, o, gelu=gelu)
    run(f"{nm} LayerNorm-fold consumer", A, W, b, o, gelu=gelu, ln=(st, c))
