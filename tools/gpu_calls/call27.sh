#!/bin/bash
# GPU call 27: launch list of the pixel decoder + mask path at 8 images (what carries the 17.2 ms), ncu --set full of the uint8 staging kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c27_launches_heads8.csv python tools/prof_heads.py 8 > gpurun_out/c27_ncu_list.log 2>&1; tail -1 gpurun_out/c27_ncu_list.log
python tools/launch_summary.py gpurun_out/c27_launches_heads8.csv | head -45
timeout 300 ncu --clock-control none --set full -k regex:stage_u8 -c 1 -o /tmp/stage -f python tools/prof_heads.py 4 > gpurun_out/c27_ncu_stage.log 2>&1
ncu -i /tmp/stage.ncu-rep --page raw --csv > gpurun_out/c27_ncu_raw_stage.csv 2>/dev/null; rm -f /tmp/stage.ncu-rep
