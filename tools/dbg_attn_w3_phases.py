"""Debug aid: per-role cycle accounting of the pipelined windowed attention kernel (attention_win3.cu, PH instantiation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iuvl_b200 import cabi
DEV = "cuda"; lib = cabi.lib(); st = cabi.stream_ptr
B, g, ws, heads, hd = int(os.environ.get("B", 16)), 64, 14, 16, 80
D = heads * hd
gen = torch.Generator().manual_seed(1)
qkv = torch.randn(B, 70, 70, 3 * D, generator=gen).bfloat16().to(DEV)
pack = torch.zeros(lib.svb_rel_pack_rows(ws, g), hd, dtype=torch.bfloat16, device=DEV)
for w in (0, 1):
    t = (torch.randn(27, hd, generator=gen) * 0.1).to(DEV)
    cabi.check(lib.svb_pack_rel_table(t.data_ptr(), pack.data_ptr(), 27, hd, w, st()), "pack")
out = torch.empty(B * g * g, D, dtype=torch.bfloat16, device=DEV)
clk = torch.zeros(64, dtype=torch.int64, device=DEV)
names = {0: ["other", "wait BIASR", "bias fetch", "wait SFULL", "softmax"], 1: ["other", "wait BIASR", "bias fetch", "wait SFULL", "softmax"],
         2: ["wait BIASF0", "skew0", "wait PV0", "epi0", "wait BIASF1", "skew1", "wait PV1", "epi1", "other"],
         3: ["loop", "wait QKFULL", "wait BIASC", "issue bias", "wait PFULL", "issue PV", "wait OREAD", "issue S"]}
for rep in range(2):
    clk.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cabi.check(lib.svb_attention_tc_phases(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, clk.data_ptr(), st()), "attn")
    e1.record(); torch.cuda.synchronize()
    c = clk.cpu().reshape(4, 16).double()
    print(f"rep {rep}: {e0.elapsed_time(e1) * 1e3:.0f} us (instrumented)")
    for r, label in ((0, "softmax group 0"), (1, "softmax group 1"), (2, "helper"), (3, "issuer 0")):
        n = max(1.0, c[r, 11].item())
        print(f"  {label}: items {n:.0f}; cycles per item:", {nm: round(c[r, k].item() / n) for k, nm in enumerate(names[r])},
              "total", round(c[r, :11].sum().item() / n))
