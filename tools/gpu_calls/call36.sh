#!/bin/bash
# GPU call 36: programmatic dependent launch extended to the attention and pad-row kernels: parity (encoder + operator suites), A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/c36_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -4 gpurun_out/c36_pytest.log | cut -c1-300
for v in 0 1 0 1; do
  SVB_PDL=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c36_bench_pdl$v.json 2> gpurun_out/c36_bench.err
  echo "SVB_PDL=$v $(python tools/summarize_bench.py gpurun_out/c36_bench_pdl$v.json | cut -c1-140)"
done | tee gpurun_out/c36_pdl_ab.txt
