mkdir -p gpurun_out
for v in 1 0 1 0; do SVB_PAD_IN_GEMM=$v timeout 600 python bench.py --batch 16 --steps 4 --no-cpu-baseline --no-e2e > gpurun_out/bench_pf$v.json 2> gpurun_out/bench_pf$v.err; python tools/summarize_bench.py gpurun_out/bench_pf$v.json | cut -c1-330; done
