#!/bin/bash
# GPU call 30: per-level operand cache of the cross-attention layers; parity of the operator / N1 / N4 suites, pipeline timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_msda.py tests/test_gpu_pixel_decoder.py tests/test_gpu_mask_head.py -m gpu -x -q > gpurun_out/c30_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -4 gpurun_out/c30_pytest.log | cut -c1-400
timeout 600 python tools/pipeline_bench.py 2>&1 | tail -2 | tee gpurun_out/c30_pipeline.txt
timeout 600 python tools/pipeline_bench.py 4 2>&1 | tail -1 | tee -a gpurun_out/c30_pipeline.txt
