mkdir -p gpurun_out
export SVB_GEMM2_VERBOSE=1
echo "== cluster 4 check"; SVB_GEMM_CLUSTER=4 timeout 120 python tools/gemm_bench.py --check --reps 5 --no-cublas 2>&1 | tail -12
echo "== cluster 4 tests"; SVB_GEMM_CLUSTER=4 timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" 2>&1 | tail -4
echo "== cluster 2"; SVB_GEMM_CLUSTER=2 timeout 120 python tools/gemm_bench.py --reps 40 2>&1 | tail -7
echo "== cluster 4"; SVB_GEMM_CLUSTER=4 timeout 120 python tools/gemm_bench.py --reps 40 --no-cublas 2>&1 | tail -7
for c in 2 4; do SVB_GEMM_CLUSTER=$c timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo rc=$?; python tools/summarize_bench.py gpurun_out/bench_c$c.json; done
