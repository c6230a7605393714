import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iuvl_b200 as ib
from iuvl_b200 import cabi
from tests.test_gpu_ops import _ref_attention_grid
DEV = "cuda"; lib = cabi.lib(); st = cabi.stream_ptr
for (gh, gw, heads, hd, B, scale_tab) in ((64, 64, 1, 64, 1, 0.0), (64, 64, 1, 64, 1, 0.2), (32, 64, 1, 80, 1, 0.0), (32, 64, 1, 80, 1, 0.2), (64, 128, 2, 80, 1, 0.2)):
    D = heads * hd
    gen = torch.Generator().manual_seed(5)
    qkv = torch.randn(B * gh * gw, 3 * D, generator=gen).bfloat16()
    rel_h, rel_w = torch.randn(2 * gh - 1, hd, generator=gen) * scale_tab, torch.randn(2 * gw - 1, hd, generator=gen) * scale_tab
    ref = _ref_attention_grid(qkv.float(), rel_h.bfloat16().float(), rel_w.bfloat16().float(), torch.zeros(3 * D), B, gh, gw, 0, heads)
    nws = int(lib.svb_attention_global_hw_workspace(B, gh, gw, heads, hd))
    ws = torch.zeros(nws + 256, dtype=torch.uint8, device=DEV)
    base = (ws.data_ptr() + 255) & ~255
    out = torch.full((B * gh * gw, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    qd, hd_, wd = qkv.to(DEV), rel_h.to(DEV), rel_w.to(DEV)
    cabi.check(lib.svb_attention_global_hw(qd.data_ptr(), out.data_ptr(), hd_.data_ptr(), wd.data_ptr(), B, gh, gw, heads, hd, base, nws, st()), "g")
    torch.cuda.synchronize()
    M = B * gh * gw
    ldh, ldw = (2 * gh - 1 + 7) // 8 * 8, (2 * gw - 1 + 7) // 8 * 8
    off = base - ws.data_ptr()
    bh = ws[off:off + M * heads * ldh * 4].view(torch.float32).reshape(M, heads, ldh).cpu()
    bwt = ws[off + M * heads * ldh * 4: off + M * heads * (ldh + ldw) * 4].view(torch.float32).reshape(M, heads, ldw).cpu()
    q = qkv.float()[:, :D].reshape(M, heads, hd)
    want_h = torch.einsum("mhc,jc->mhj", q, rel_h.bfloat16().float())
    want_w = torch.einsum("mhc,jc->mhj", q, rel_w.bfloat16().float())
    eh = float((bh[:, :, :2 * gh - 1] - want_h).abs().max()); ew = float((bwt[:, :, :2 * gw - 1] - want_w).abs().max())
    print(f"grid {gh}x{gw} heads {heads} hd {hd} tab {scale_tab}: rel_l2 {ib.rel_l2(out, ref):.3e}; bias table max abs err h {eh:.3e} w {ew:.3e} (scale {float(want_h.abs().max()):.2f})")
    o3, r3 = out.float().cpu().reshape(B, gh, gw, D), ref.float().reshape(B, gh, gw, D)
    print("   per-row-band rel_l2:", [round(ib.rel_l2(o3[:, y:y + 8], r3[:, y:y + 8]), 3) for y in range(0, gh, 8)][:8])
