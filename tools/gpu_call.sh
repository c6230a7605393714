mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_encoder.py -q -x -k "attention or tiny_bf16 or full_size" 2>&1 | tail -3
SVB_PROF_DETAIL=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python tools/summarize_bench.py gpurun_out/bench_default.json
grep "prof cat 4" gpurun_out/bench_default.err | sort -k12 -n -r | head -4
