mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msda.py -q -x 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -x -k "uint8" 2>&1 | tail -5
timeout 300 python tools/msda_bench.py 8 2>&1 | tail -3
python tools/prof_step.py --batch 8 --steps 2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc2s -s 150 -c 8 -o gpurun_out/prof_gemm python tools/prof_step.py --batch 8 --steps 2 > gpurun_out/ncu_full.log 2>&1; echo full rc=$?; tail -2 gpurun_out/ncu_full.log
