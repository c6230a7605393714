mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" 2>&1 | tail -3
timeout 300 python tools/parity_report.py vit_h_std bf16 2>&1 | tail -1
timeout 300 python tools/parity_report.py vit_h_stress bf16 2>&1 | tail -1
for v in 1 0 1 0; do SVB_ATTNG_ONEPASS=$v timeout 600 python bench.py --batch 16 --steps 4 --no-cpu-baseline --no-e2e > gpurun_out/bench_op$v.json 2> gpurun_out/bench_op$v.err; python tools/summarize_bench.py gpurun_out/bench_op$v.json | cut -c1-330; done
