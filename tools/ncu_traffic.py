#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (read HERE, no GPU needed): per captured launch the duration, DRAM bytes, tensor-pipe
and L2 utilisation -> a CSV under profiles/<dir>/ and profiles/traffic.json (the per-launch DRAM traffic bench.py reports).

    python tools/ncu_traffic.py gpurun_out/prof_gemm.ncu-rep profiles/r01_call40 vit_h
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, outdir, model = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "vit_h")
os.makedirs(outdir, exist_ok=True)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [head.index(w) for w in want if w in head]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
with open(os.path.join(outdir, "ncu_full_selected_metrics.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([head[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i][:80] for i in idx])
tot, n = 0.0, 0
ir, iw, ik = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum"), head.index("Kernel Name")
for r in data:
    if "gemm_tc2" in r[ik]:
        tot += float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0)
        n += 1
tj = os.path.join(os.path.dirname(os.path.abspath(outdir.rstrip("/"))), "traffic.json")
d = {}
if os.path.exists(tj):
    d = json.load(open(tj))
d[model] = {"gemm_bytes_per_launch_avg": tot / n if n else None, "gemm_launches_captured": n, "capture": os.path.basename(rep),
            "summary": os.path.join(os.path.basename(outdir.rstrip("/")), "ncu_full_selected_metrics.csv")}
json.dump(d, open(tj, "w"), indent=1)
print(json.dumps(d[model]))
