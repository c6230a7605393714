timeout 600 python -m pytest tests/test_gpu_msda.py -q -x 2>&1 | tail -6
timeout 300 python tools/msda_module_bench.py 8 2>&1 | tail -3
