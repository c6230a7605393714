#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "other_token_grids" > gpurun_out/c8_pytest_ops.log 2>&1; echo "ops exit $?"; grep -E "passed|failed|FAILED|AssertionError" gpurun_out/c8_pytest_ops.log | cut -c1-300
