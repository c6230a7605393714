mkdir -p gpurun_out
for c in 16 32 24 16 32; do timeout 600 python bench.py --chunk $c --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ck$c.json 2> gpurun_out/bench_ck$c.err; echo "chunk $c"; python tools/summarize_bench.py gpurun_out/bench_ck$c.json | cut -c1-300; done
