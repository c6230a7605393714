#!/usr/bin/env python
"""One GEMM variant a few times (ncu target).  python tools/one_gemm.py proj_resid|proj_plain|lin2_fused|lin1_fold [reps]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200 import cabi  # noqa: E402

dev = "cuda"
lib = cabi.lib()
D, M = 1280, 8 * 4096
which = sys.argv[1] if len(sys.argv) > 1 else "proj_resid"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
parts = (D + 127) // 128
X = torch.randn(M, D, device=dev)
Xb = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
st = torch.zeros(M, parts, 2, device=dev)
st[:, :, 1] = 128.0
cfgs = {
    "proj_plain": (D, D, False, False, False, False),
    "proj_resid": (D, D, True, False, False, False),
    "proj_fused": (D, D, True, True, False, False),
    "lin2_resid": (D, 4 * D, True, False, False, False),
    "lin2_fused": (D, 4 * D, True, True, False, False),
    "lin1_plain": (4 * D, D, False, False, False, True),
    "lin1_fold": (4 * D, D, False, False, True, True),
}
n, k, resid, extras, fold, gelu = cfgs[which]
A = torch.randn(M, k, device=dev).bfloat16()
W = (torch.randn(n, k, device=dev) / math.sqrt(k)).bfloat16()
b = torch.randn(n, device=dev)
c = torch.randn(n, device=dev)
out = X if resid else torch.empty(M, n, dtype=torch.bfloat16, device=dev)
for _ in range(reps):
    rc = lib.svb_linear_fused(A.data_ptr(), k, W.data_ptr(), k, M, n, k, b.data_ptr(), int(gelu), X.data_ptr() if resid else None,
                              D if resid else 0, 0, out.data_ptr(), cabi.DTYPE_F32 if resid else cabi.DTYPE_BF16, n,
                              st.data_ptr() if fold else None, c.data_ptr() if fold else None, D if fold else 0, 1e-6,
                              Xb.data_ptr() if extras else None, D if extras else 0, st.data_ptr() if extras else None, 0, 0,
                              cabi.stream_ptr())
    cabi.check(rc, which)
torch.cuda.synchronize()
print("ok", which)
