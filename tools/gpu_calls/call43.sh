#!/bin/bash
# GPU call 43: ncu --set full of the streaming kernels added in the last session (HBM GB/s against the measured peak)
mkdir -p gpurun_out
PH="python tools/prof_heads.py 8"
timeout 900 ncu --clock-control none --set full -k regex:"layernorm_post|fpn_gn_up_pad|resize_aa_cols|resize_aa_rows|mask_threshold_clear|stage_u8" -c 14 -o /tmp/newk -f $PH > gpurun_out/c43_ncu.log 2>&1; tail -2 gpurun_out/c43_ncu.log
ncu -i /tmp/newk.ncu-rep --page raw --csv > gpurun_out/c43_ncu_raw_newkernels.csv 2>/dev/null; rm -f /tmp/newk.ncu-rep
python tools/ncu_raw_summary.py gpurun_out/c43_ncu_raw_newkernels.csv > gpurun_out/ncu_full_newkernels.csv 2>/dev/null; head -c 600 gpurun_out/ncu_full_newkernels.csv
