#!/bin/bash
# GPU call 3: pipelined windowed kernel (attention_win3.cu): parity, then A/B timings of the kernel variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention or softmax" > gpurun_out/c3_pytest_ops.log 2>&1; echo "ops exit $?"; tail -15 gpurun_out/c3_pytest_ops.log | cut -c1-300
for v in "SVB_ATTNW_IMPL=2" "SVB_ATTNW_IMPL=3" "SVB_ATTNW_IMPL=3 SVB_ATTNW_POLY=0" "SVB_ATTNW_IMPL=3 SVB_ATTNW_POLY=3" "SVB_ATTNW_IMPL=3 SVB_ATTNW_POLY=4" "SVB_ATTNG_POLY=2" "SVB_ATTNG_POLY=3" "SVB_ATTNG_POLY=4"; do
  env $v timeout 300 python tools/attn_bench.py 2>&1 | tail -1
done | tee gpurun_out/c3_attn_ab.txt
HD=64 HEADS=12 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | tee -a gpurun_out/c3_attn_ab.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "benchmarked or tiny_bf16 or full_size or full_tensor" > gpurun_out/c3_pytest_enc.log 2>&1; echo "enc exit $?"; tail -5 gpurun_out/c3_pytest_enc.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/c3_bench.json | cut -c1-500
