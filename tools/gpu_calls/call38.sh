#!/bin/bash
# GPU call 38: exposure-aware pass schedule of the host path: tests, then A/B of the end-to-end number (SVB_HOST_SCHEDULE)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -s -k "schedule or host_path or chunking" > gpurun_out/c38_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; grep -a "host schedule" gpurun_out/c38_pytest.log; tail -3 gpurun_out/c38_pytest.log | cut -c1-300
for v in 0 1 0 1; do
  SVB_HOST_SCHEDULE=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/c38_bench_hs$v.json 2> gpurun_out/c38_bench.err
  echo "SVB_HOST_SCHEDULE=$v $(python tools/summarize_bench.py gpurun_out/c38_bench_hs$v.json | sed 's/(ms, TF.*| e2e/| e2e/' | cut -c1-200)"
done | tee gpurun_out/c38_host_schedule_ab.txt
