#!/bin/bash
# GPU call 13: staggered windowed attention kernel (SVB_ATTNW_IMPL=4): parity of every variant, then kernel-alone timings
mkdir -p gpurun_out
T="tests/test_gpu_ops.py -m gpu -x -q -k"
for v in "4 3 1" "4 3 0" "4 0 1" "4 0 0"; do
  set -- $v
  SVB_ATTNW_IMPL=$1 SVB_ATTNW_MODE=$2 SVB_ATTNW_SPARE=$3 timeout 600 python -m pytest $T "test_attention_tcgen05 and not variants" > gpurun_out/c13_pytest_$1_$2_$3.log 2>&1
  echo "impl $1 mode $2 spare $3: pytest exit $?"; tail -3 gpurun_out/c13_pytest_$1_$2_$3.log | cut -c1-300
done
for rep in 1 2; do
for v in "2 0 0" "4 0 0" "4 1 0" "4 2 0" "4 3 0" "4 0 1" "4 1 1" "4 3 1" "4 7 1" "4 4 0"; do
  set -- $v
  SVB_ATTNW_IMPL=$1 SVB_ATTNW_MODE=$2 SVB_ATTNW_SPARE=$3 timeout 300 python tools/attn_bench.py 2>&1 | tail -1
done
done | tee gpurun_out/c13_attn_ab.txt
for v in "4 3 1" "4 3 0"; do
  set -- $v
  SVB_ATTNW_POLY=0 SVB_ATTNW_IMPL=$1 SVB_ATTNW_MODE=$2 SVB_ATTNW_SPARE=$3 timeout 300 python tools/attn_bench.py 2>&1 | tail -1
  SVB_ATTNW_L2AHEAD=2 SVB_ATTNW_IMPL=$1 SVB_ATTNW_MODE=$2 SVB_ATTNW_SPARE=$3 timeout 300 python tools/attn_bench.py 2>&1 | tail -1
  SVB_ATTNW_L2AHEAD=0 SVB_ATTNW_IMPL=$1 SVB_ATTNW_MODE=$2 SVB_ATTNW_SPARE=$3 timeout 300 python tools/attn_bench.py 2>&1 | tail -1
done | tee -a gpurun_out/c13_attn_ab.txt
HD=64 HEADS=12 SVB_ATTNW_IMPL=4 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | tee -a gpurun_out/c13_attn_ab.txt
HD=64 HEADS=12 SVB_ATTNW_IMPL=2 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | tee -a gpurun_out/c13_attn_ab.txt
