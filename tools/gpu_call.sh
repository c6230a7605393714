mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -x -k "uint8" 2>&1 | tail -3
echo "== bf16 two boxes 5 stages"; timeout 300 python tools/gemm_bench_fused.py 10 6 "LayerNorm-fold" 2>&1 | tail -2
echo "== bf16 one box 6 stages"; SVB_GEMM_BF16_ONEBOX=1 timeout 300 python tools/gemm_bench_fused.py 10 6 "LayerNorm-fold" 2>&1 | tail -2
echo "== bf16 two boxes 5 stages"; timeout 300 python tools/gemm_bench_fused.py 10 6 "LayerNorm-fold" 2>&1 | tail -2
echo "== bf16 one box 6 stages"; SVB_GEMM_BF16_ONEBOX=1 timeout 300 python tools/gemm_bench_fused.py 10 6 "LayerNorm-fold" 2>&1 | tail -2
timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1]); print(d['next_rows']['stage_u8'])"
for v in 0 1; do SVB_GEMM_BF16_ONEBOX=$v timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ob$v.json 2> gpurun_out/bench_ob$v.err; python tools/summarize_bench.py gpurun_out/bench_ob$v.json | cut -c1-330; done
