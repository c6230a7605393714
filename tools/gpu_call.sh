mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/t_all.log 2>&1; echo all rc=$?; tail -3 gpurun_out/t_all.log
echo "== no prefetch"; timeout 300 python tools/gemm_bench_fused.py 10 6 "fold producer" 2>&1 | grep -E "proj|lin2"
echo "== L2 prefetch"; SVB_GEMM2_DBG=16 timeout 300 python tools/gemm_bench_fused.py 10 6 "fold producer" 2>&1 | grep -E "proj|lin2"
timeout 300 python tools/parity_report.py vit_h_std bf16 2>&1 | tail -1
timeout 300 python tools/parity_report.py vit_h_stress bf16 2>&1 | tail -1
timeout 300 python tools/parity_report.py vit_l_std bf16 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
python tools/summarize_bench.py gpurun_out/bench_default.json
