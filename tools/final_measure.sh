# Round-end measurement pass (run under gpurun): bench (both arms), launch list, ncu --set full captures of the top kernels.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python tools/summarize_bench.py gpurun_out/final_bench.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; cut -c1-200 gpurun_out/final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches2.csv python tools/prof_step.py --batch 16 --steps 1 > gpurun_out/ncu_list2.log 2>&1; tail -1 gpurun_out/ncu_list2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_window_persistent -s 2 -c 2 -o gpurun_out/prof_attnw_final -f python tools/prof_step.py --batch 16 --steps 1 > gpurun_out/ncu_w.log 2>&1; tail -1 gpurun_out/ncu_w.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_global2 -c 2 -o gpurun_out/prof_attng_final -f python tools/prof_step.py --batch 16 --steps 1 > gpurun_out/ncu_g.log 2>&1; tail -1 gpurun_out/ncu_g.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2s -s 8 -c 8 -o gpurun_out/prof_gemm_final -f python tools/prof_step.py --batch 16 --steps 1 > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
