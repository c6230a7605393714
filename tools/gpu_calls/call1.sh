#!/bin/bash
# GPU call 1 of round 2: full GPU test suite + the bench with every extra record.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/c1_bench.json 2>/dev/null | cut -c1-400
tail -3 gpurun_out/c1_bench.err
