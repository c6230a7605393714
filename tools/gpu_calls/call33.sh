#!/bin/bash
# GPU call 33: launch lists of ViT-B batch 16 and ViT-L batch 16 passes (BASELINE configs 2 / 3: where the step time goes)
mkdir -p gpurun_out
for m in vit_b vit_l; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c33_launches_${m}_b16.csv python tools/prof_step.py --model $m --batch 16 --steps 2 > gpurun_out/c33_ncu_$m.log 2>&1; tail -1 gpurun_out/c33_ncu_$m.log
  python tools/launch_summary.py gpurun_out/c33_launches_${m}_b16.csv 2>/dev/null | head -14
done
