"""tcgen05.mma cost per shape / operand source on one SM (svb_probe_mma_rate): cycles per MMA, back to back."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iuvl_b200 import cabi
lib = cabi.lib()
names = {0: "SS N=128 K-major B (floor 64)", 1: "SS N=64 K-major B (floor 32)", 2: "TS N=64 K-major B (32)", 3: "TS N=64 MN-major B (32)",
         4: "TS N=16 MN-major SW32 (8)", 5: "PV step: TS N=64 + N=16 MN-major (40)", 6: "TS N=128 K-major B (64)", 7: "SS N=208 K-major B (104)",
         8: "TS N=64 + N=16 K-major (40)", 9: "TS N=80 MN-major SW128, two atoms (40)", 10: "TS N=80 MN-major SW32, five atoms (40)"}
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for alt in (0,):
    for v, nm in names.items():
        if v in (0, 1, 4, 7, 8, 2, 6): continue
        res = []
        for reps in (64, 512):
            for _ in range(2):
                cabi.check_probe(cabi.probe_lib().svb_probe_mma_rate(v, reps, alt, out.data_ptr(), cabi.stream_ptr()), "rate")
                torch.cuda.synchronize()
            res.append(out.cpu().tolist())
        per = (res[1][0] - res[0][0]) / (512 - 64)
        print(f"flags={alt} (2: commit every 4, 4: background tcgen05.ld/st) {nm:46s} cycles/MMA {per:7.1f}   (total {res[1][0]}, issue loop {res[1][1]} for 512)")
