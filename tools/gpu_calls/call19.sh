#!/bin/bash
# GPU call 19: TMEM load / store rate of one SM (probe) + the windowed kernel's two softmax passes with their TMEM loads knocked out
mkdir -p gpurun_out
timeout 300 python tools/tmem_rate.py 2>&1 | tee gpurun_out/c19_tmem_rate.txt
for rep in 1 2; do
for ko in 0 512 1536 2560 3584 3592 513 514 1024 2048 3072; do
  SVB_ATTNW_KO=$ko timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
done
done | tee gpurun_out/c19_ko.txt
