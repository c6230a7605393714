#!/bin/bash
# GPU call 35: programmatic dependent launch of the GEMMs (SVB_PDL): parity of the encoder suite with it on, then A/B of the step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/c35_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -4 gpurun_out/c35_pytest.log | cut -c1-300
for v in 0 1 0 1 0 1; do
  SVB_PDL=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c35_bench_pdl$v.json 2> gpurun_out/c35_bench.err
  echo "SVB_PDL=$v $(python tools/summarize_bench.py gpurun_out/c35_bench_pdl$v.json | cut -c1-140)"
done | tee gpurun_out/c35_pdl_ab.txt
