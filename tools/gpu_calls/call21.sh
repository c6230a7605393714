#!/bin/bash
# GPU call 21: round-end state: the whole GPU suite, smoke(), the bench with every record and the reference arm, launch lists at the
# pass sizes the bench runs, ncu --set full captures of the streaming kernels and of the N1 / N4 kernels (summarised on the box)
mkdir -p gpurun_out /tmp/ncu
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/c21_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/c21_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c21_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/c21_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/c21_bench.json 2> gpurun_out/c21_bench.err; echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/c21_bench.json | cut -c1-600
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c21_ref.json 2> gpurun_out/c21_ref.err; echo "reference arm exit $?"; cut -c1-400 gpurun_out/c21_ref.json
NCU="ncu --clock-control none"
P16="python tools/prof_step.py --batch 16 --steps 1"
P12="python tools/prof_step.py --batch 12 --steps 1"
PH="python tools/prof_heads.py 4"
$P16 > gpurun_out/c21_plain16.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/r02f_launches_b16.csv $P16 > gpurun_out/c21_ncu1.log 2>&1; tail -1 gpurun_out/c21_ncu1.log
$P12 > gpurun_out/c21_plain12.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/r02f_launches_b12.csv $P12 > gpurun_out/c21_ncu2.log 2>&1; tail -1 gpurun_out/c21_ncu2.log
full() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 sk=$3 cnt=$4; shift 4
  timeout 900 $NCU --set full -k regex:"$rx" -s $sk -c $cnt -o /tmp/ncu/$name -f "$@" > gpurun_out/c21_ncu_$name.log 2>&1
  tail -1 gpurun_out/c21_ncu_$name.log
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/r02f_ncu_raw_$name.csv 2>/dev/null
}
full stream "gn_apply|cast_s2d|im2col_kernel|fill_pad|stage_u8" 0 10 $P16
$PH > gpurun_out/c21_plainh.log 2>&1 && timeout 600 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file gpurun_out/r02f_launches_heads.csv $PH > gpurun_out/c21_ncu7.log 2>&1; tail -1 gpurun_out/c21_ncu7.log
full heads "xattn_tc_kernel|msda_fused|resize_aa|pad_rows_bf16|gn_rows|nchw_to_rows|rows_to_nchw|upsample_add|mask_threshold|cls_token" 150 80 $PH
rm -rf /tmp/ncu
du -sh gpurun_out
