#!/bin/bash
# GPU call 25: shared-arithmetic MSDA kernels (fused + plain): parity, the whole rows N1 / N4 suites, then timings and the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_msda.py tests/test_gpu_pixel_decoder.py tests/test_gpu_mask_head.py -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -6 gpurun_out/c25_pytest.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
for sh in 0 1; do
  echo "SVB_MSDA_SHARED=$sh"
  SVB_MSDA_SHARED=$sh timeout 300 python tools/msda_bench.py 2>&1 | tail -3
done | tee gpurun_out/c25_msda_ab.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err; echo "bench exit $?"
python tools/summarize_bench.py gpurun_out/c25_bench.json | cut -c1-600
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c25_bench.json').read().strip().splitlines()[-1])
nr = d['next_rows']
for k in ('msda_forward', 'msda_module', 'pixel_decoder', 'xdecoder_mask_path'):
    print(k, {a: b for a, b in nr[k].items() if a != 'workload'})
p = nr['pipeline']; print('pipeline', p['value'], p['stage_ms_rank0'], p['e2e']['value'])
PY
PH="python tools/prof_heads.py 4"
timeout 600 ncu --clock-control none --set full -k regex:msda_fused_shared -s 1 -c 1 -o /tmp/msda -f $PH > gpurun_out/c25_ncu_msda.log 2>&1
ncu -i /tmp/msda.ncu-rep --page raw --csv > gpurun_out/r02f_ncu_raw_heads_msda_shared.csv 2>/dev/null; rm -f /tmp/msda.ncu-rep
