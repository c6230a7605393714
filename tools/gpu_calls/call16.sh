#!/bin/bash
# GPU call 16: knock-out timing of the two-group windowed kernel with L2-resident operands (KO bit 512: every item = window 0 of
# image 0), i.e. the SM-side cost structure without the HBM floor
mkdir -p gpurun_out
for rep in 1 2; do
for ko in 0 512 513 514 515 516 520 528 544 576 640 768 960 1012 523 527 543 548 564; do
  SVB_ATTNW_KO=$ko timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
done
done | tee gpurun_out/c16_ko.txt
