#!/bin/bash
# GPU call 7: scope row N3 on the tensor cores (other token grids): op-level parity, encoder goldens, timing against the fp32-math path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "other_token_grids or attention_tcgen05" > gpurun_out/c7_pytest_ops.log 2>&1; echo "ops exit $?"; tail -12 gpurun_out/c7_pytest_ops.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "other_canvases or tiny_bf16 or resize" > gpurun_out/c7_pytest_enc.log 2>&1; echo "enc exit $?"; tail -12 gpurun_out/c7_pytest_enc.log | cut -c1-400
for v in 1 0; do SVB_ATTN_EXT=$v timeout 600 python tools/canvas_bench.py vit_b 1024 2048 4 2>&1 | tail -2; done | tee gpurun_out/c7_canvas.txt
SVB_ATTN_EXT=1 timeout 600 python tools/canvas_bench.py vit_h 1024 2048 2 2>&1 | tail -2 | tee -a gpurun_out/c7_canvas.txt
