#!/bin/bash
# GPU call 10: new GPU tests (step1 geometry, other token grids), then the ncu evidence of round 2 (launch lists at the pass sizes the
# bench runs — 16 and 12 images — and --set full captures of the kernels this round touched)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pixel_decoder.py tests/test_gpu_mask_head.py -m gpu -x -q -k "other_token_grids or step1_geometry" > gpurun_out/c10_pytest.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/c10_pytest.log | cut -c1-300
P16="python tools/prof_step.py --batch 16 --steps 1"
P12="python tools/prof_step.py --batch 12 --steps 1"
NCU="ncu --clock-control none"
$P16 > gpurun_out/c10_plain16.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/r02_launches_b16.csv $P16 > gpurun_out/c10_ncu1.log 2>&1; tail -1 gpurun_out/c10_ncu1.log
$P12 > gpurun_out/c10_plain12.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/r02_launches_b12.csv $P12 > gpurun_out/c10_ncu2.log 2>&1; tail -1 gpurun_out/c10_ncu2.log
$P16 > gpurun_out/c10_plain16b.log 2>&1 && timeout 600 $NCU --set full --import-source on -k regex:attn_global2 -c 1 -o gpurun_out/r02_prof_attng -f $P16 > gpurun_out/c10_ncu3.log 2>&1; tail -1 gpurun_out/c10_ncu3.log
$P16 > gpurun_out/c10_plain16c.log 2>&1 && timeout 600 $NCU --set full --import-source on -k regex:attn_window_persistent -s 2 -c 1 -o gpurun_out/r02_prof_attnw -f $P16 > gpurun_out/c10_ncu4.log 2>&1; tail -1 gpurun_out/c10_ncu4.log
$P16 > gpurun_out/c10_plain16d.log 2>&1 && timeout 600 $NCU --set full --import-source on -k regex:gemm_tc2s -s 8 -c 8 -o gpurun_out/r02_prof_gemm -f $P16 > gpurun_out/c10_ncu5.log 2>&1; tail -1 gpurun_out/c10_ncu5.log
$P16 > gpurun_out/c10_plain16e.log 2>&1 && timeout 600 $NCU --set full -k regex:"gn_apply|cast_s2d|im2col_kernel|fill_pad" -c 8 -o gpurun_out/r02_prof_stream -f $P16 > gpurun_out/c10_ncu6.log 2>&1; tail -1 gpurun_out/c10_ncu6.log
PH="python tools/prof_heads.py 4"
$PH > gpurun_out/c10_plainh.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -s 0 -c 3000 --csv --log-file gpurun_out/r02_launches_heads.csv $PH > gpurun_out/c10_ncu7.log 2>&1; tail -1 gpurun_out/c10_ncu7.log
$PH > gpurun_out/c10_plainh2.log 2>&1 && timeout 900 $NCU --set full --import-source on -k regex:"xattn_tc_kernel|msda_fused|resize_aa|im2col3x3|gn_rows|nchw_to_rows|rows_to_nchw|upsample_add|stage_u8|mask_threshold|cls_token" -s 200 -c 60 -o gpurun_out/r02_prof_heads -f $PH > gpurun_out/c10_ncu8.log 2>&1; tail -1 gpurun_out/c10_ncu8.log
PC="python tools/canvas_bench.py vit_b 1024 2048 1"
$PC > gpurun_out/c10_plainc.log 2>&1 && timeout 600 $NCU --set full --import-source on -k regex:attn_global_ext -c 1 -o gpurun_out/r02_prof_ext -f $PC > gpurun_out/c10_ncu9.log 2>&1; tail -1 gpurun_out/c10_ncu9.log
ls -la gpurun_out/*.ncu-rep
