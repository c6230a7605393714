"""Kernel-alone timing of the two tcgen05 attention kernels at the ViT-H geometry (CUDA events, B images per launch); the variant is
chosen by the environment (SVB_ATTNW_IMPL, SVB_ATTNW_POLY, SVB_ATTNG_POLY), read once per process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iuvl_b200 import cabi
DEV = "cuda"; lib = cabi.lib(); st = cabi.stream_ptr
B, g, heads, hd = int(os.environ.get("B", 16)), 64, int(os.environ.get("HEADS", 16)), int(os.environ.get("HD", 80))
D = heads * hd
gen = torch.Generator().manual_seed(1)
res = {}
for ws in (14, 64):
    if ws == 14:
        qkv = torch.randn(B, 70, 70, 3 * D, generator=gen).bfloat16().to(DEV)
    else:
        qkv = torch.randn(B * g * g, 3 * D, generator=gen).bfloat16().to(DEV)
    rows = lib.svb_rel_pack_rows(ws, g)
    pack = torch.zeros(rows, hd, dtype=torch.bfloat16, device=DEV)
    L = 2 * ws - 1
    rh, rw = (torch.randn(L, hd, generator=gen) * 0.1).to(DEV), (torch.randn(L, hd, generator=gen) * 0.1).to(DEV)
    cabi.check(lib.svb_pack_rel_table(rh.data_ptr(), pack.data_ptr(), L, hd, 0, st()), "pack")
    cabi.check(lib.svb_pack_rel_table(rw.data_ptr(), pack.data_ptr(), L, hd, 1, st()), "pack")
    out = torch.empty(B * g * g, D, dtype=torch.bfloat16, device=DEV)
    for _ in range(3):
        cabi.check(lib.svb_attention_tc(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, st()), "attn")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        cabi.check(lib.svb_attention_tc(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, st()), "attn")
    e1.record(); torch.cuda.synchronize()
    res[ws] = e0.elapsed_time(e1) * 1e3 / n
env = {k: v for k, v in os.environ.items() if k.startswith("SVB_")}
print(f"B={B} hd={hd} heads={heads} {env}: windowed {res[14]:.1f} us, global {res[64]:.1f} us per launch")
