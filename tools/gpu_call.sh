mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" > gpurun_out/t_attn.log 2>&1; echo attn rc=$?; tail -5 gpurun_out/t_attn.log
timeout 200 python tools/dbg_attn_phases.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_encoder.py -q -x > gpurun_out/t_enc.log 2>&1; echo enc rc=$?; tail -5 gpurun_out/t_enc.log
timeout 200 python tools/dbg_chunk.py tiny80 15 2>&1 | tail -3
timeout 600 python bench.py --batch 16 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_w2.json 2> gpurun_out/bench_w2.err; echo b2 rc=$?
python tools/summarize_bench.py gpurun_out/bench_w2.json
