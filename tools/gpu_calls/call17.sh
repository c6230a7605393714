#!/bin/bash
# GPU call 17: which TMA traffic carries the floor of the windowed kernel (chains kernel, diagnostic flags in SVB_ATTNW_L2AHEAD >> 8)
mkdir -p gpurun_out
for rep in 1 2; do
for f in 1 0 257 513 1025 2049 769 1793 3841 3840; do
  SVB_ATTNW_IMPL=5 SVB_ATTNW_L2AHEAD=$f timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
done
done | tee gpurun_out/c17_tma.txt
