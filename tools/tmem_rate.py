"""TMEM load / store rate of one SM (svb_probe_tmem_rate): bytes per clock for 1..16 warps issuing 32x32b.x32 loads / stores back to back."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iuvl_b200 import cabi
lib = cabi.probe_lib(); st = cabi.stream_ptr
out = torch.zeros(16, dtype=torch.int64, device="cuda")
reps = 2000
for mode, name in ((0, "ld, one in flight"), (1, "ld, two in flight"), (2, "st"), (3, "ld 16x256b.x8"), (4, "ld 16x128b.x16"), (5, "ld 16x64b.x32")):
    for nw in (1, 2, 4, 8, 12, 16):
        out.zero_()
        cabi.check_probe(lib.svb_probe_tmem_rate(nw, reps, mode, out.data_ptr(), st()), "tmem_rate")
        torch.cuda.synchronize()
        cyc = out[:nw].tolist()
        print(f"{name:18s} warps {nw:2d}: {max(cyc) / reps:7.1f} cycles per x32 op and warp, {nw * reps * 4096 / max(cyc):7.1f} B/clk/SM")
