#!/bin/bash
# GPU call 12: the whole GPU suite on the current build, the bench with every record, ncu capture of the N1 / N2 / N4 kernels
mkdir -p gpurun_out /tmp/ncu
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c12_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/c12_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/c12_bench.json | cut -c1-500
PH="python tools/prof_heads.py 4"
$PH > gpurun_out/c12_plainh.log 2>&1 && timeout 900 ncu --clock-control none --set full -k regex:"xattn_tc_kernel|msda_fused|resize_aa|pad_rows_bf16|gn_rows|nchw_to_rows|rows_to_nchw|upsample_add|stage_u8|mask_threshold|cls_token" -s 150 -c 80 -o /tmp/ncu/heads -f $PH > gpurun_out/c12_ncu_heads.log 2>&1; tail -1 gpurun_out/c12_ncu_heads.log
ncu -i /tmp/ncu/heads.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_heads.csv 2>/dev/null; ls -la /tmp/ncu/
rm -rf /tmp/ncu
