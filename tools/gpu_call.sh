mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/t_ops.log 2>&1; echo ops rc=$?; tail -5 gpurun_out/t_ops.log
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_probe.py -q -x > gpurun_out/t_enc.log 2>&1; echo enc rc=$?; tail -5 gpurun_out/t_enc.log
timeout 200 python tools/dbg_chunk.py tiny64 25 2>&1 | tail -3
timeout 200 python tools/dbg_chunk.py tiny80 15 2>&1 | tail -3
timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_b16.json 2> gpurun_out/bench_b16.err; echo b2 rc=$?
python tools/summarize_bench.py gpurun_out/bench_b16.json
timeout 300 python tools/prof_step.py --steps 2 > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 300 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --steps 2 > gpurun_out/ncu_list.log 2>&1; echo list rc=$?
