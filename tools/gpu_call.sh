mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" 2>&1 | tail -6
echo "== streamlined"; timeout 120 python tools/gemm_bench.py --reps 40 --no-cublas --check 2>&1 | tail -9
echo "== generic"; SVB_GEMM_EPI=0 timeout 120 python tools/gemm_bench.py --reps 40 --no-cublas 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_encoder.py -q -x 2>&1 | tail -4
timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo rc=$?; python tools/summarize_bench.py gpurun_out/bench_s.json
SVB_LN_FOLD=1 timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_sf.json 2> gpurun_out/bench_sf.err; echo rc=$?; python tools/summarize_bench.py gpurun_out/bench_sf.json
