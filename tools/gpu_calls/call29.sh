#!/bin/bash
# GPU call 29: transposed-store GEMM for the mask-logit einsum (svb_linear_nt), width fallback of the fused post-norm pass;
# parity of rows N1 / N4 + the new operator tests, then timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_msda.py tests/test_gpu_pixel_decoder.py tests/test_gpu_mask_head.py -m gpu -x -q > gpurun_out/c29_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -8 gpurun_out/c29_pytest.log | cut -c1-400
timeout 300 python tools/pixel_decoder_bench.py 8 2>&1 | tail -2 | tee gpurun_out/c29_pixdec8.txt
timeout 300 python tools/mask_head_bench.py 8 2>&1 | head -3 | tee gpurun_out/c29_mask_head.txt
timeout 600 python tools/pipeline_bench.py 2>&1 | tail -2 | tee gpurun_out/c29_pipeline.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c29_launches_heads8.csv python tools/prof_heads.py 8 > gpurun_out/c29_ncu_list.log 2>&1; tail -1 gpurun_out/c29_ncu_list.log
python tools/launch_summary.py gpurun_out/c29_launches_heads8.csv 2>/dev/null | head -32
