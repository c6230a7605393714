#!/bin/bash
# GPU call 11: the ncu evidence of round 2, summarised ON THE BOX (gpurun_out/ is limited to 64 MiB: the .ncu-rep files are converted to
# their raw-page CSV — and, for the global attention kernel, the source page — and removed)
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
P16="python tools/prof_step.py --batch 16 --steps 1"
P12="python tools/prof_step.py --batch 12 --steps 1"
PH="python tools/prof_heads.py 4"
PC="python tools/canvas_bench.py vit_b 1024 2048 1"
full() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 sk=$3 cnt=$4; shift 4
  "$@" > gpurun_out/c11_plain_$name.log 2>&1 && timeout 900 $NCU --set full --import-source on -k regex:"$rx" -s $sk -c $cnt -o /tmp/ncu/$name -f "$@" > gpurun_out/c11_ncu_$name.log 2>&1
  tail -1 gpurun_out/c11_ncu_$name.log
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_$name.csv 2>/dev/null
  ls -la /tmp/ncu/$name.ncu-rep
}
timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "uint8" > gpurun_out/c11_pytest.log 2>&1; echo "uint8 tests exit $?"; tail -2 gpurun_out/c11_pytest.log
$P16 > gpurun_out/c11_plain16.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/r02_launches_b16.csv $P16 > gpurun_out/c11_ncu1.log 2>&1; tail -1 gpurun_out/c11_ncu1.log
$P12 > gpurun_out/c11_plain12.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/r02_launches_b12.csv $P12 > gpurun_out/c11_ncu2.log 2>&1; tail -1 gpurun_out/c11_ncu2.log
full attng "attn_global2" 0 1 $P16
ncu -i /tmp/ncu/attng.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_source_attng.csv.gz
full attnw "attn_window_persistent" 2 1 $P16
ncu -i /tmp/ncu/attnw.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_source_attnw.csv.gz
full gemm "gemm_tc2s" 8 8 $P16
full stream "gn_apply|cast_s2d|im2col_kernel|fill_pad" 0 8 $P16
$PH > gpurun_out/c11_plainh.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file gpurun_out/r02_launches_heads.csv $PH > gpurun_out/c11_ncu7.log 2>&1; tail -1 gpurun_out/c11_ncu7.log
full heads "xattn_tc_kernel|msda_fused|resize_aa|im2col3x3|gn_rows|nchw_to_rows|rows_to_nchw|upsample_add|stage_u8|mask_threshold|cls_token" 200 60 $PH
full ext "attn_global_ext" 0 1 $PC
rm -rf /tmp/ncu
du -sh gpurun_out
