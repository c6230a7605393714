#!/bin/bash
# GPU call 44: fused FPN kernel with 32-bit index arithmetic: parity of the pixel decoder suite, timing at 8 / 16 images
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pixel_decoder.py -m gpu -x -q > gpurun_out/c44_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -3 gpurun_out/c44_pytest.log | cut -c1-300
for n in 8 16 8; do timeout 300 python tools/pixel_decoder_bench.py $n 2>&1 | tail -1; done | tee gpurun_out/c44_pixdec.txt
