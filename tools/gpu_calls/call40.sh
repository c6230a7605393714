#!/bin/bash
# GPU call 40: threshold + clear-full-rows in one kernel: parity of the mask head suite, A/B of the mask path, pipeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mask_head.py -m gpu -x -q > gpurun_out/c40_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -4 gpurun_out/c40_pytest.log | cut -c1-300
for f in 0 1 0 1; do echo "SVB_MASK_CLEAR_FUSE=$f $(SVB_MASK_CLEAR_FUSE=$f timeout 600 python tools/pipeline_bench.py 8 2>&1 | tail -1)"; done | tee gpurun_out/c40_mask_clear_ab.txt
