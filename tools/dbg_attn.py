"""Debug aid: run-to-run determinism of the tcgen05 attention kernels; prints where mismatching outputs sit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iuvl_b200 import cabi
DEV = "cuda"
lib = cabi.lib()
st = lambda: cabi.stream_ptr()

def run(qkv_src, pack, B, g, ws, heads, hd):
    out = torch.full((B * g * g, heads * hd), float("nan"), dtype=torch.bfloat16, device=DEV)
    cabi.check(lib.svb_attention_tc(qkv_src.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, st()), "attn")
    torch.cuda.synchronize()
    return out

for ws, heads, hd, B in [(14, 16, 80, 4), (14, 12, 64, 4)]:
    g = 64; D = heads * hd
    gen = torch.Generator().manual_seed(ws + hd)
    qkv = torch.randn(B * g * g, 3 * D, generator=gen).bfloat16().to(DEV)
    L = 2 * ws - 1
    rel_h = (torch.randn(L, hd, generator=gen) * 0.1).to(DEV); rel_w = (torch.randn(L, hd, generator=gen) * 0.1).to(DEV)
    bias = torch.randn(3 * D, generator=gen).to(DEV)
    pack = torch.zeros(lib.svb_rel_pack_rows(ws, g), hd, dtype=torch.bfloat16, device=DEV)
    lib.svb_pack_rel_table(rel_h.data_ptr(), pack.data_ptr(), L, hd, 0, st()); lib.svb_pack_rel_table(rel_w.data_ptr(), pack.data_ptr(), L, hd, 1, st())
    if ws != g:
        gp = 70
        src = torch.zeros(B, gp, gp, 3 * D, dtype=torch.bfloat16, device=DEV)
        src[:, :g, :g] = qkv.reshape(B, g, g, 3 * D)
        lib.svb_fill_pad_rows(src.data_ptr(), bias.data_ptr(), B, g, gp, 3 * D, st())
    else:
        src = qkv
    ref = run(src, pack, B, g, ws, heads, hd)
    nbad = 0
    for rep in range(40):
        o = run(src, pack, B, g, ws, heads, hd)
        diff = (o.float() != ref.float()) | (o.float().isnan() != ref.float().isnan())
        if diff.any():
            nbad += 1
            if nbad <= 3:
                idx = diff.nonzero()
                rows = idx[:, 0].unique()
                b = rows // (g * g); y = (rows % (g * g)) // g; x = rows % g
                cols = idx[:, 1].unique()
                print(f"  ws={ws} hd={hd} rep={rep}: {len(rows)} rows differ; images {b.unique().tolist()}; y {y.min().item()}..{y.max().item()} x {x.min().item()}..{x.max().item()};"
                      f" window(s) {sorted(set(zip((y // ws).tolist(), (x // ws).tolist())))[:6]}; heads {(cols // hd).unique().tolist()}; cols-in-head {(cols % hd).min().item()}..{(cols % hd).max().item()};"
                      f" maxabs {float((o.float() - ref.float())[diff].abs().max()):.3e}")
    print(f"ws={ws} heads={heads} hd={hd} B={B}: {nbad}/40 runs differ from the first")
