#!/bin/bash
# GPU call 23: ncu --set full of the row N1 / N4 kernels (one launch of each), summarised on the box
mkdir -p gpurun_out /tmp/ncu
PH="python tools/prof_heads.py 4"
$PH > gpurun_out/c23_plainh.log 2>&1; tail -1 gpurun_out/c23_plainh.log
i=0
for k in msda_fused xattn_tc_kernel xattn_tc_combine pad_rows_bf16 gn_rows_stats gn_rows_apply nchw_to_rows rows_to_nchw upsample_add add_cast_kernel add_cast_bcast layernorm_fixed nchw_to_seq mask_clear_full_rows; do
  i=$((i+1))
  timeout 600 ncu --clock-control none --set full -k regex:"$k" -s 1 -c 1 -o /tmp/ncu/h$i -f $PH > gpurun_out/c23_ncu_$k.log 2>&1
  ncu -i /tmp/ncu/h$i.ncu-rep --page raw --csv > gpurun_out/r02f_ncu_raw_heads_$k.csv 2>/dev/null
  echo "$k: $(tail -1 gpurun_out/c23_ncu_$k.log | cut -c1-80)"
done
# the implicit-GEMM 3x3 convolution: the GEMM launch that follows pad_rows_bf16
timeout 600 ncu --clock-control none --set full -k regex:"pad_rows_bf16|gemm_tc2s" -c 200 -o /tmp/ncu/conv -f $PH > gpurun_out/c23_ncu_conv.log 2>&1
ncu -i /tmp/ncu/conv.ncu-rep --page raw --csv > gpurun_out/r02f_ncu_raw_heads_conv.csv 2>/dev/null
rm -rf /tmp/ncu; du -sh gpurun_out
