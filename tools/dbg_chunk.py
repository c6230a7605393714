"""Debug aid: repeated chunked forwards of a tiny encoder; counts runs that differ from the first run of the same chunking."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iuvl_b200 as ib
from iuvl_b200.encoder import build_encoder
name = sys.argv[1] if len(sys.argv) > 1 else "tiny64"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = ib.PRESETS[name]
sd = ib.make_state_dict(cfg, 2024, rel_std=0.1)
x = ib.make_images(2, cfg, 3)
x3 = torch.cat([x, x[:1]], 0).cuda()
enc = build_encoder(cfg); enc.load_state_dict(sd); enc.cuda()
enc.enable_taps(True) if os.environ.get("TAPS") else None
first, bad = {}, {}
for rep in range(reps):
    for chunk in (1, 2, 3):
        enc.max_chunk = chunk
        with torch.no_grad():
            o = enc(x3)
        torch.cuda.synchronize()
        if chunk not in first:
            first[chunk] = o
            continue
        errs = [max(float(ib.rel_l2(o[k][i], first[chunk][k][i])) for k in o) for i in range(3)]
        if max(errs) > 0:
            bad.setdefault(chunk, []).append((rep, ["%.1e" % e for e in errs]))
ref = first[3]
for c in (1, 2):
    print("chunk", c, "vs 3 (first runs):", [max(float(ib.rel_l2(first[c][k][i], ref[k][i])) for k in ref) for i in range(3)])
print("nondeterministic runs:", {c: len(v) for c, v in bad.items()}, "of", reps - 1, "| samples:", {c: v[:3] for c, v in bad.items()})
