#!/bin/bash
# GPU call 2: warp-uniform MMA issue in the attention kernels: parity, phase clocks, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention or softmax" > gpurun_out/c2_pytest_ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/c2_pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "benchmarked or tiny_bf16 or full_size" > gpurun_out/c2_pytest_enc.log 2>&1; echo "enc exit $?"; tail -3 gpurun_out/c2_pytest_enc.log
timeout 300 python tools/dbg_attn_g_phases.py > gpurun_out/c2_phases_g.txt 2>&1; cat gpurun_out/c2_phases_g.txt
timeout 300 python tools/dbg_attn_phases.py > gpurun_out/c2_phases_w.txt 2>&1; cat gpurun_out/c2_phases_w.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/c2_bench.json | cut -c1-500
