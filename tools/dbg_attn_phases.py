"""Debug aid: per-phase cycle breakdown of the persistent windowed attention kernel's softmax warp groups."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iuvl_b200 import cabi
DEV = "cuda"; lib = cabi.lib(); st = cabi.stream_ptr
B, g, ws, heads, hd = 8, 64, 14, 16, 80
D = heads * hd
gen = torch.Generator().manual_seed(1)
qkv = torch.randn(B, 70, 70, 3 * D, generator=gen).bfloat16().to(DEV)
pack = (torch.randn(64, hd, generator=gen) * 0.1).bfloat16().to(DEV)
out = torch.empty(B * g * g, D, dtype=torch.bfloat16, device=DEV)
clk = torch.zeros(16, dtype=torch.int64, device=DEV)
for rep in range(3):
    clk.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cabi.check(lib.svb_attention_tc_phases(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, clk.data_ptr(), st()), "attn")
    e1.record(); torch.cuda.synchronize()
    c = clk.cpu().reshape(2, 8).double()
    names = ["wait bias", "bias read+shift", "wait S", "softmax", "wait PV", "epilogue"]
    print(f"rep {rep}: {e0.elapsed_time(e1) * 1e3:.0f} us")
    for i in range(2):
        n = c[i, 6].item()
        print(f"  group {i}: items {n:.0f}; cycles per item:", {nm: round(c[i, k].item() / n) for k, nm in enumerate(names)}, "total", round(c[i, :6].sum().item() / n))
# plain timing without counters
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    cabi.check(lib.svb_attention_tc(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, st()), "attn")
e1.record(); torch.cuda.synchronize()
print("plain: %.0f us per launch" % (e0.elapsed_time(e1) * 1e3 / 5))
