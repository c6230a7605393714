#!/bin/bash
# GPU call 6: warp-uniform MMA issue in the GEMM: parity + interleaved A/B (SVB_GEMM2_DBG=32 = former single-thread issue), then the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm or linear or layernorm" > gpurun_out/c6_pytest_ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/c6_pytest_ops.log | cut -c1-300
for r in 1 2 3; do
  for v in "SVB_GEMM2_DBG=0" "SVB_GEMM2_DBG=32"; do
    echo "== $v"; env $v timeout 300 python tools/gemm_bench.py --images 16 --reps 40 --no-cublas 2>&1 | tail -5
  done
done | tee gpurun_out/c6_gemm_ab.txt
for v in "SVB_GEMM2_DBG=0" "SVB_GEMM2_DBG=32" "SVB_GEMM2_DBG=0" "SVB_GEMM2_DBG=32"; do
  env $v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/c6_bench_$v.json 2> gpurun_out/c6_bench.err
  echo "$v"; python tools/summarize_bench.py gpurun_out/c6_bench_$v.json | cut -c1-330
done | tee gpurun_out/c6_bench_ab.txt
