#!/bin/bash
# GPU call 20: production windowed kernel = two chains + one-pass softmax: parity (incl. the adversarial case), kernel-alone and in-step timing
mkdir -p gpurun_out
T="tests/test_gpu_ops.py -m gpu -x -q -k"
timeout 900 python -m pytest $T "attention or softmax" > gpurun_out/c20_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -12 gpurun_out/c20_pytest.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
for rep in 1 2; do
  for v in "2 1" "5 1" "5 0"; do set -- $v; SVB_ATTNW_IMPL=$1 SVB_ATTNW_ONEPASS=$2 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'; done
  SVB_ATTNW_POLY=0 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
  SVB_ATTNW_L2AHEAD=3841 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
  SVB_ATTNW_L2AHEAD=2 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
  for impl in 2 5; do B=12 SVB_ATTNW_IMPL=$impl timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'; done
  for impl in 2 5; do HD=64 HEADS=12 SVB_ATTNW_IMPL=$impl timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'; done
done | tee gpurun_out/c20_attn_ab.txt
for impl in 2 5 2 5; do
  SVB_ATTNW_IMPL=$impl timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/c20_bench_$impl.json 2> gpurun_out/c20_bench_$impl.err; echo "bench impl $impl exit $?"
  python tools/summarize_bench.py gpurun_out/c20_bench_$impl.json | cut -c1-420
done
