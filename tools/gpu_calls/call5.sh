#!/bin/bash
# GPU call 5: windowed kernels with L2 prefetch / non-blocking early bias: parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention or softmax" > gpurun_out/c5_pytest_ops.log 2>&1; echo "ops exit $?"; tail -5 gpurun_out/c5_pytest_ops.log | cut -c1-300
for v in "SVB_ATTNW_IMPL=2 SVB_ATTNW_L2AHEAD=0" "SVB_ATTNW_IMPL=2 SVB_ATTNW_L2AHEAD=1" "SVB_ATTNW_IMPL=2 SVB_ATTNW_L2AHEAD=2" "SVB_ATTNW_IMPL=3 SVB_ATTNW_L2AHEAD=0" "SVB_ATTNW_IMPL=3 SVB_ATTNW_L2AHEAD=1" "SVB_ATTNW_IMPL=3 SVB_ATTNW_L2AHEAD=2" "SVB_ATTNW_IMPL=3 SVB_ATTNW_L2AHEAD=3"; do
  env $v timeout 300 python tools/attn_bench.py 2>&1 | tail -1
done | tee gpurun_out/c5_attn_ab.txt
SVB_ATTNW_L2AHEAD=2 timeout 300 python tools/dbg_attn_w3_phases.py > gpurun_out/c5_w3_phases.txt 2>&1; tail -5 gpurun_out/c5_w3_phases.txt
