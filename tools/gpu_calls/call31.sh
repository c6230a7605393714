#!/bin/bash
# GPU call 31: FPN level fused around the implicit 3x3 convolution (svb_fpn_conv3x3_rows), bf16 operand of mask_features straight from
# GroupNorm + ReLU; parity of the pixel decoder / chain tests, A/B timing at 8 images
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pixel_decoder.py tests/test_gpu_mask_head.py -m gpu -x -q > gpurun_out/c31_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -4 gpurun_out/c31_pytest.log | cut -c1-400
for f in 0 1 0 1; do echo "SVB_FPN_FUSE=$f"; SVB_FPN_FUSE=$f timeout 300 python tools/pixel_decoder_bench.py 8 2>&1 | tail -1; done | tee gpurun_out/c31_pixdec8.txt
