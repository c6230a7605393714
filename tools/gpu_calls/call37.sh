#!/bin/bash
# GPU call 37: which dependent launches pay: SVB_PDL mask 0 / 1 (GEMMs) / 3 (+ attention) / 7 (+ pad rows), alternating
mkdir -p gpurun_out
for v in 0 1 3 7 0 1 3 7; do
  SVB_PDL=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c37_bench_pdl$v.json 2> gpurun_out/c37_bench.err
  echo "SVB_PDL=$v $(python tools/summarize_bench.py gpurun_out/c37_bench_pdl$v.json | cut -c1-110)"
done | tee gpurun_out/c37_pdl_ab.txt
