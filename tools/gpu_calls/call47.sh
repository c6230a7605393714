#!/bin/bash
# GPU call 47: fraction of the global kernel's exponentials on the FMA pipe (SVB_ATTNG_POLY = pairs of 8) inside the ViT-B and ViT-H steps
mkdir -p gpurun_out
for m in "vit_b 16" "vit_h 64"; do
  set -- $m
  for v in 2 0 3 4 2 0; do
    SVB_ATTNG_POLY=$v timeout 600 python bench.py --model $1 --batch $2 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c47_bench.json 2> gpurun_out/c47_bench.err
    echo "$1 SVB_ATTNG_POLY=$v $(python tools/summarize_bench.py gpurun_out/c47_bench.json | sed 's/.*json //' | cut -c1-60) attn_global $(python -c "import json;d=json.loads(open('gpurun_out/c47_bench.json').read().strip().splitlines()[-1]);print(round(d['roofline']['categories']['attn_global']['ms_per_step'],2))")"
  done
done | tee gpurun_out/c47_attng_poly.txt
