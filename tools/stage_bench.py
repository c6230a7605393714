#!/usr/bin/env python
"""Timing of the uint8 input staging kernel (row N2: svb_stage_images_u8) against the measured HBM peak.

    python tools/stage_bench.py [sets]

`sets` independent groups of 16 images / destination buffers are visited round-robin (default 4: 600 MB in flight, well
beyond the 126 MB L2), so the number is an HBM number and not an L2 one; `sets = 1` reproduces bench.py's next_rows.stage_u8."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iuvl_b200 import cabi  # noqa: E402

sets = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev, B = "cuda", 16
lib = cabi.lib()
peak = 6457.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
groups = []
for s in range(sets):
    imgs = [torch.randint(0, 256, (3, 1024, 1024), dtype=torch.uint8, device=dev) for _ in range(B)]
    dst = torch.empty(B * 4096, 768, dtype=torch.bfloat16, device=dev)
    groups.append((imgs, dst, (C.c_void_p * B)(*[t.data_ptr() for t in imgs])))
hs, ws = (C.c_int * B)(*([1024] * B)), (C.c_int * B)(*([1024] * B))
mean, std = (C.c_float * 3)(123.675, 116.28, 103.53), (C.c_float * 3)(58.395, 57.12, 57.375)


def run(g):
    cabi.check(lib.svb_stage_images_u8(g[2], hs, ws, B, 3, 1024, 16, mean, std, g[1].data_ptr(), cabi.DTYPE_BF16, cabi.stream_ptr()))


for g in groups:
    run(g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for r in range(reps):
    for g in groups:
        run(g)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (reps * sets)
byts = B * 3 * 1024 * 1024 * (1 + 2)
print(f"stage_u8 ({sets} x 16 images round-robin): {ms * 1e3:.1f} us per launch, {byts / ms / 1e6:.0f} GB/s = "
      f"{byts / ms / 1e6 / peak:.3f} of the measured HBM peak ({peak:.0f} GB/s)")
# bit-exactness of one group against the expression the reference's caller evaluates ((x - mean) / std in fp32, then bf16)
imgs, dst, _ = groups[0]
m = torch.tensor([123.675, 116.28, 103.53], device=dev).view(3, 1, 1)
s = torch.tensor([58.395, 57.12, 57.375], device=dev).view(3, 1, 1)
ref = ((imgs[3].float() - m) / s).view(3, 64, 16, 64, 16).permute(1, 3, 0, 2, 4).reshape(4096, 768).bfloat16()
assert torch.equal(ref, dst[3 * 4096:4 * 4096]), "stage_u8 output differs from the fp32 expression"
print("stage_u8 bit-exact")
