#!/bin/bash
# GPU call 14: knock-out timing of the two-group windowed kernel (diagnostic build, SVB_BUILD_KO=1): which phase carries the item time
mkdir -p gpurun_out
for rep in 1 2; do
for ko in 0 1 2 3 4 8 16 32 64 128 256 448 36 52 500 11 15 31; do
  SVB_ATTNW_KO=$ko timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
done
done | tee gpurun_out/c14_ko.txt
python tools/dbg_attn_phases.py 2>&1 | tail -12 | tee gpurun_out/c14_phases.txt
