mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -x 2>&1 | tail -3
for c in 16 8; do timeout 600 python bench.py --chunk $c --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench_chunk$c.err; echo rc=$?; python tools/summarize_bench.py gpurun_out/bench_chunk$c.json; done
python tools/prof_step.py --batch 8 --steps 2 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --batch 8 --steps 2 > gpurun_out/ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2s -s 400 -c 8 -o gpurun_out/prof_gemm python tools/prof_step.py --batch 8 --steps 2 > gpurun_out/ncu_full.log 2>&1; echo full rc=$?; tail -2 gpurun_out/ncu_full.log
