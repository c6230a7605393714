#!/usr/bin/env python
"""Interleaved A/B timing of the GEMM epilogue variants (svb_linear_fused), ViT-H shapes, 8 images (SVB_BENCH_IMAGES=n: n images).  The variants are run
round-robin (R rounds of `reps` launches each) so that every one sees the same clock / power state of the capped GPU."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200 import cabi  # noqa: E402

dev = "cuda"
lib = cabi.lib()
D, M = 1280, int(os.environ.get("SVB_BENCH_IMAGES", "8")) * 4096
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
only = sys.argv[3] if len(sys.argv) > 3 else ""
parts = (D + 127) // 128


def mk(m, n, k):
    A = torch.randn(m, k, device=dev).bfloat16()
    W = (torch.randn(n, k, device=dev) / math.sqrt(k)).bfloat16()
    return A, W, torch.randn(n, device=dev)


X = torch.randn(M, D, device=dev)
Xb = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
st = torch.zeros(M, parts, 2, device=dev)
st[:, :, 1] = 128.0
variants = []


def add(name, A, W, bias, out, gelu=False, resid=None, ln=None, out2=None, stat=None):
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
    if only in name:
        variants.append([name, call, 2.0 * Mx * N * K, 0.0])


for nm, k in (("proj", D), ("lin2", 4 * D)):
    A, W, b = mk(M, D, k)
    ob = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    add(f"{nm} bf16 out", A, W, b, ob)
    add(f"{nm} fp32 out, no residual (TMA store)", A, W, b, X)
    add(f"{nm} fp32 in-place residual (TMA reduce-add)", A, W, b, X, resid=X)
    add(f"{nm} fold producer (+bf16 copy +row sums)", A, W, b, X, resid=X, out2=Xb, stat=st)
for nm, n, gelu in (("qkv", 3 * D, False), ("lin1", 4 * D, True)):
    A, W, b = mk(M, n, D)
    c = torch.randn(n, device=dev)
    o = torch.empty(M, n, dtype=torch.bfloat16, device=dev)
    add(f"{nm} plain", A, W, b, o, gelu=gelu)
    add(f"{nm} LayerNorm-fold consumer", A, W, b, o, gelu=gelu, ln=(st, c))

for v in variants:
    v[1]()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(rounds):
    for v in variants:
        e0.record()
        for _ in range(reps):
            v[1]()
        e1.record()
        torch.cuda.synchronize()
        if r > 0:          # round 0 = warm-up
            v[3] += e0.elapsed_time(e1) / reps
for name, _, fl, ms in variants:
    ms /= max(1, rounds - 1)
    print(f"{name:50s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TF/s", flush=True)
