timeout 600 python -m pytest tests/test_gpu_msda.py -q -x 2>&1 | tail -3
timeout 300 python tools/msda_bench.py 4 2>&1 | tail -2
timeout 300 python tools/msda_bench.py 8 2>&1 | tail -2
