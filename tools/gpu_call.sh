mkdir -p gpurun_out
python tools/prof_step.py --batch 12 --steps 2 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --batch 12 --steps 2 > gpurun_out/ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2s -s 140 -c 8 -o gpurun_out/prof_gemm python tools/prof_step.py --batch 12 --steps 2 > gpurun_out/ncu_full.log 2>&1; echo full rc=$?; tail -2 gpurun_out/ncu_full.log
