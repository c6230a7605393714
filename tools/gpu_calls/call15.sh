#!/bin/bash
# GPU call 15: two-independent-chains windowed kernel (SVB_ATTNW_IMPL=5): parity, kernel-alone timing, in-step timing
mkdir -p gpurun_out
T="tests/test_gpu_ops.py -m gpu -x -q -k"
SVB_ATTNW_IMPL=5 timeout 600 python -m pytest $T "test_attention_tcgen05 and not variants" > gpurun_out/c15_pytest.log 2>&1
echo "impl 5: pytest exit $?"; tail -3 gpurun_out/c15_pytest.log | cut -c1-300
for rep in 1 2; do
  for impl in 2 5; do SVB_ATTNW_IMPL=$impl timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'; done
  SVB_ATTNW_IMPL=5 SVB_ATTNW_POLY=0 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
  SVB_ATTNW_IMPL=5 SVB_ATTNW_L2AHEAD=0 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
  SVB_ATTNW_IMPL=5 SVB_ATTNW_L2AHEAD=2 timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'
  for impl in 2 5; do B=12 SVB_ATTNW_IMPL=$impl timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'; done
  for impl in 2 5; do HD=64 HEADS=12 SVB_ATTNW_IMPL=$impl timeout 300 python tools/attn_bench.py 2>&1 | tail -1 | sed 's/, global.*//'; done
done | tee gpurun_out/c15_attn_ab.txt
for impl in 2 5 2 5; do
  SVB_ATTNW_IMPL=$impl timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/c15_bench_$impl.json 2> gpurun_out/c15_bench_$impl.err; echo "bench impl $impl exit $?"
  python tools/summarize_bench.py gpurun_out/c15_bench_$impl.json | cut -c1-420
done
