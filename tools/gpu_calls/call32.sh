#!/bin/bash
# GPU call 32: the whole GPU suite, smoke, the default bench (timed), the reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c32_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -8 gpurun_out/c32_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c32_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/c32_smoke.log
( time timeout 1500 python bench.py > gpurun_out/c32_bench.json 2> gpurun_out/c32_bench.err ) 2> gpurun_out/c32_bench.time; echo "bench exit $?"; cat gpurun_out/c32_bench.time
python tools/summarize_bench.py gpurun_out/c32_bench.json | cut -c1-900
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c32_ref.json 2> gpurun_out/c32_ref.err ) 2> gpurun_out/c32_ref.time; cut -c1-300 gpurun_out/c32_ref.json; cat gpurun_out/c32_ref.time
