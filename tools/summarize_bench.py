#!/usr/bin/env python
"""Print the per-category timing of bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        r = d["roofline"]
        cats = {k: (round(v["ms_per_step"], 2), v["tflops"] and round(v["tflops"]), v["gbs"] and round(v["gbs"]))
                for k, v in r["categories"].items()}
        print(path, round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), "| whole-step frac of sustained",
              round(r["whole_step"]["frac_of_sustained"], 3), "| (ms, TF/s, GB/s):", cats, "| e2e", d.get("e2e") and round(d["e2e"]["value"], 1),
              "| clocks", d.get("clocks"))
    except Exception as e:  # noqa: BLE001
        print(path, "ERR", e)
