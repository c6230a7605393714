#!/bin/bash
# GPU call 39: exposure-aware host schedule after the small-chunk fix: tests, end-to-end A/B incl. the ViT-B / ViT-L records
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -s -k "schedule or host_path or chunking" > gpurun_out/c39_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; grep -a "host schedule" gpurun_out/c39_pytest.log; tail -2 gpurun_out/c39_pytest.log | cut -c1-300
for v in 0 1; do
  SVB_HOST_SCHEDULE=$v timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c39_bench_hs$v.json 2> gpurun_out/c39_bench.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c39_bench_hs$v.json').read().strip().splitlines()[-1])
print('SVB_HOST_SCHEDULE=$v', 'vit_h value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), {k:(round(c['value'],1), round(c['e2e'],1)) for k,c in d['configs'].items()}, 'pipeline', round(d['next_rows']['pipeline']['value'],1), round(d['next_rows']['pipeline']['e2e']['value'],1))
PY
done | tee gpurun_out/c39_host_schedule_ab.txt
