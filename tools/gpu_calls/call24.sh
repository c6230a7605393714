#!/bin/bash
# GPU call 24: MSDA fused kernel with the per-unit arithmetic shared by the unit's lanes: parity, then A/B timing (module, pixel decoder)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msda.py tests/test_gpu_pixel_decoder.py -m gpu -x -q > gpurun_out/c24_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -6 gpurun_out/c24_pytest.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
for rep in 1 2; do
  for sh in 0 1; do
    echo "SVB_MSDA_SHARED=$sh"
    SVB_MSDA_SHARED=$sh timeout 300 python tools/msda_module_bench.py 4 2>&1 | tail -2
    SVB_MSDA_SHARED=$sh timeout 300 python tools/pixel_decoder_bench.py 4 2>&1 | tail -2
  done
done | tee gpurun_out/c24_msda_ab.txt
