#!/bin/bash
# GPU call 49: pad rows written by idle warps of the qkv GEMM (SVB_PAD_IN_GEMM=1) re-measured now that the GEMM -> attention chain
# uses dependent launches
mkdir -p gpurun_out
for v in 0 1 0 1; do
  SVB_PAD_IN_GEMM=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c49_bench.json 2> gpurun_out/c49_bench.err
  echo "SVB_PAD_IN_GEMM=$v $(python tools/summarize_bench.py gpurun_out/c49_bench.json | sed 's/.*json //' | cut -c1-150)"
done | tee gpurun_out/c49_pad_in_gemm.txt
