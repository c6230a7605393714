#!/bin/bash
# GPU call 4: phase accounting of the pipelined windowed kernel; tcgen05 cross-attention + class/box heads parity and timing
mkdir -p gpurun_out
timeout 300 python tools/dbg_attn_w3_phases.py > gpurun_out/c4_w3_phases.txt 2>&1; cat gpurun_out/c4_w3_phases.txt
timeout 900 python -m pytest tests/test_gpu_mask_head.py -m gpu -x -q > gpurun_out/c4_pytest_mh.log 2>&1; echo "mask head exit $?"; tail -15 gpurun_out/c4_pytest_mh.log | cut -c1-300
timeout 300 python tools/mask_head_bench.py > gpurun_out/c4_mask_head_bench.txt 2>&1; tail -12 gpurun_out/c4_mask_head_bench.txt
SVB_XATTN_IMPL=0 timeout 300 python tools/mask_head_bench.py > gpurun_out/c4_mask_head_bench_simt.txt 2>&1; tail -12 gpurun_out/c4_mask_head_bench_simt.txt
