mkdir -p gpurun_out
timeout 300 python tools/parity_report.py vit_h_std bf16 2>&1 | tail -1
SVB_GN_FOLD=0 timeout 300 python tools/parity_report.py vit_h_std bf16 2>&1 | tail -1
timeout 300 python tools/parity_report.py vit_b_std bf16 2>&1 | tail -1
timeout 300 python tools/parity_report.py tiny80_stress bf16 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ops.py -q -x 2>&1 | tail -3
for v in 1 0 1 0; do SVB_PROF_DETAIL=1 SVB_GN_FOLD=$v timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_gn$v.json 2> gpurun_out/bench_gn$v.err; python tools/summarize_bench.py gpurun_out/bench_gn$v.json | cut -c1-330; done
grep "prof cat 0" gpurun_out/bench_gn1.err | sort -k12 -n -r | sed -n 5,20p
