#!/bin/bash
# GPU call 42: state at the end of round 2: the whole GPU suite, smoke, the default bench (timed), the reference arm, launch lists
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c42_pytest.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -8 gpurun_out/c42_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c42_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/c42_smoke.log
( time timeout 1500 python bench.py > gpurun_out/c42_bench.json 2> gpurun_out/c42_bench.err ) 2> gpurun_out/c42_bench.time; echo "bench exit $?"; cat gpurun_out/c42_bench.time
python tools/summarize_bench.py gpurun_out/c42_bench.json | cut -c1-900
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c42_bench.json').read().strip().splitlines()[-1])
print({k: (round(c['value'], 1), round(c['e2e'], 1), round(c['whole_encoder_frac'], 3)) for k, c in d['configs'].items()})
nr = d['next_rows']
for k in ('pixel_decoder', 'xdecoder_mask_path', 'stage_u8'):
    print(k, {a: b for a, b in nr[k].items() if a != 'workload'})
p = nr['pipeline']; print('pipeline', round(p['value'], 1), p['stage_ms_rank0'], round(p['e2e']['value'], 1))
PY
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c42_ref.json 2> gpurun_out/c42_ref.err ) 2> gpurun_out/c42_ref.time; cut -c1-200 gpurun_out/c42_ref.json
for b in 12 16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02h_launches_b$b.csv python tools/prof_step.py --batch $b --steps 1 > gpurun_out/c42_ncu_list_b$b.log 2>&1; tail -1 gpurun_out/c42_ncu_list_b$b.log
done
