#!/bin/bash
# GPU call 22 (2 GPUs): the bench under torchrun as the driver launches it: weak headline + strong leg + reference arm
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c22_bench2.json 2> gpurun_out/c22_bench2.err; echo "bench 2 GPUs exit $?"
python tools/summarize_bench.py gpurun_out/c22_bench2.json | cut -c1-600
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c22_bench2.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'n_gpus', 'scaling', 'ms_per_step')}, 'strong', d.get('strong'), 'e2e', d.get('e2e', {}).get('value'))
print('pipeline', d.get('next_rows', {}).get('pipeline', {}).get('value'), d.get('next_rows', {}).get('pipeline', {}).get('global_batch'))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/c22_ref2.json 2> gpurun_out/c22_ref2.err; echo "reference arm 2 GPUs exit $?"; cut -c1-300 gpurun_out/c22_ref2.json
tail -3 gpurun_out/c22_bench2.err
