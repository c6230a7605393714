"""Debug aid: per-phase cycle breakdown of the global attention kernel (softmax groups + MMA issuer), and ring-depth A/B timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iuvl_b200 import cabi
DEV = "cuda"; lib = cabi.lib(); st = cabi.stream_ptr
B, g, ws, heads, hd = int(os.environ.get("B", 8)), 64, 64, 16, 80
D = heads * hd
gen = torch.Generator().manual_seed(1)
qkv = torch.randn(B * g * g, 3 * D, generator=gen).bfloat16().to(DEV)
pack = (torch.randn(272, hd, generator=gen) * 0.1).bfloat16().to(DEV)
out = torch.empty(B * g * g, D, dtype=torch.bfloat16, device=DEV)
clk = torch.zeros(16, dtype=torch.int64, device=DEV)
nct = B * heads * 16
for rep in range(2):
    clk.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cabi.check(lib.svb_attention_tc_phases(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, clk.data_ptr(), st()), "attn")
    e1.record(); torch.cuda.synchronize()
    c = clk.cpu().reshape(2, 8).double()
    print(f"rep {rep}: {e0.elapsed_time(e1) * 1e3:.0f} us (instrumented)")
    for i in range(2):
        n = c[i, 6].item()
        print(f"  softmax group {i}: cycles per key tile:", {nm: round(c[i, k].item() / n) for k, nm in enumerate(["wait S", "softmax", "rescale"])})
    n = nct * 32
    print("  group 0 (half-tile kernel only): st drain + hand-over", round(c[0, 7].item() / n), " ld wait", round(c[1, 7].item() / n))
    print("  issuer: cycles per key tile:", {"wait K/V": round(c[0, 3].item() / n), "wait P (both tiles)": round(c[0, 4].item() / n), "issue": round(c[0, 5].item() / n)})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    cabi.check(lib.svb_attention_tc(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, st()), "attn")
e0.record()
for _ in range(5):
    cabi.check(lib.svb_attention_tc(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, st()), "attn")
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 5
print("plain: %.0f us per launch (NST=%s); %.0f cycles per key-tile step at 1.9 GHz x CTA slots" % (us, os.environ.get("SVB_ATTNG_NST", "3"), us * 1e-6 * 1.9e9 * 148 / nct / 32))
