#!/bin/bash
# GPU call 26: A/B of the LayerNorm-fold producer epilogue (proj / lin2) variants, uint8 staging with 16-byte stores, ncu of the
# streaming kernels (HBM GB/s evidence), pixel decoder at 8 images
mkdir -p gpurun_out
# parity of the changed code first: the interleaved-chunk variant through the fold tests, the staging kernel bit-exact
SVB_GEMM2_DBG=64 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fold or gemm or linear" > gpurun_out/c26_pytest_ilv.log 2>&1
echo "pytest (interleaved) exit $?"; tail -3 gpurun_out/c26_pytest_ilv.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_encoder.py -m gpu -x -q -k "fold or uint8 or staging or u8" > gpurun_out/c26_pytest.log 2>&1
echo "pytest (default) exit $?"; tail -3 gpurun_out/c26_pytest.log | cut -c1-300
for n in 1 4; do timeout 120 python tools/stage_bench.py $n 2>&1 | tail -2; done | tee gpurun_out/c26_stage.txt
for img in 16 12; do
  for v in "SVB_GEMM2_DBG=0" "SVB_GEMM2_DBG=64" "SVB_GEMM_RMW_ONEBOX=0" "SVB_GEMM_RMW_ONEBOX=0 SVB_GEMM2_DBG=64" "SVB_GEMM2_DBG=16" "SVB_GEMM2_DBG=80" "SVB_GEMM_RMW_ONEBOX=2 SVB_GEMM2_DBG=64" "SVB_GEMM2_DBG=0"; do
    echo "== images $img  $v"
    env $v SVB_BENCH_IMAGES=$img timeout 200 python tools/gemm_bench_fused.py 10 5 "fold producer" 2>&1 | tail -2
  done
done | tee gpurun_out/c26_gemm_ab.txt
timeout 300 python tools/pixel_decoder_bench.py 8 2>&1 | tail -4 | tee gpurun_out/c26_pixdec8.txt
