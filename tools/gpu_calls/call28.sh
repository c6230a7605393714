#!/bin/bash
# GPU call 28: deformable encoder layer with the one-pass post-norm LayerNorm (+ next layer's operands) and the in-place output
# projection; table-driven antialiased resize.  Parity of rows N1 / N4, then timings (A/B of the resize against the per-tap kernels).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_msda.py tests/test_gpu_pixel_decoder.py tests/test_gpu_mask_head.py -m gpu -x -q > gpurun_out/c28_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -8 gpurun_out/c28_pytest.log | cut -c1-400
timeout 300 python tools/pixel_decoder_bench.py 8 2>&1 | tail -2 | tee gpurun_out/c28_pixdec8.txt
for leg in 1 0; do
  echo "SVB_RESIZE_LEGACY=$leg"
  SVB_RESIZE_LEGACY=$leg timeout 300 python tools/mask_head_bench.py 8 2>&1 | head -3
done | tee gpurun_out/c28_mask_head.txt
timeout 600 python tools/pipeline_bench.py 2>&1 | tail -4 | tee gpurun_out/c28_pipeline.txt
