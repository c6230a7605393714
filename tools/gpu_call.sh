mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -x -k "canvases or resize" 2>&1 | tail -8
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/t_all.log 2>&1; echo all rc=$?; tail -4 gpurun_out/t_all.log
timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_n3.json 2> gpurun_out/bench_n3.err; echo rc=$?; python tools/summarize_bench.py gpurun_out/bench_n3.json
