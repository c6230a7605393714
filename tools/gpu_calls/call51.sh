#!/bin/bash
# GPU call 51: the last build of round 2: whole GPU suite, smoke, the default bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c51_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -6 gpurun_out/c51_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c51_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/c51_smoke.log
( time timeout 1500 python bench.py > gpurun_out/c51_bench.json 2> gpurun_out/c51_bench.err ) 2> gpurun_out/c51_bench.time; echo "bench exit $?"; tail -3 gpurun_out/c51_bench.time
python tools/summarize_bench.py gpurun_out/c51_bench.json | cut -c1-700
