#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and one block's sequence."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, seq = None, []
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].replace("void ", "").replace("unnamed>::", "").replace("svb::", "")
        seq.append((name[:34], d["Grid Size"], float(d["Metric Value"].replace(",", "")) / 1000))
agg = collections.OrderedDict()
for k, g, v in seq:
    a = agg.setdefault((k, g), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("launches", len(seq), "total %.2f ms" % (tot / 1000))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-36s %-16s n=%-4d avg %8.1f us  total %7.2f ms  %5.1f %%" % (k[0], k[1], n, t / n, t / 1000, 100 * t / tot))
if len(sys.argv) > 2:
    i0 = [i for i, (k, g, v) in enumerate(seq) if k.startswith("layernorm")][0]
    for k, g, v in seq[i0:i0 + int(sys.argv[2])]:
        print("   ", k, g, round(v, 1))
