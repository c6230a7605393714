#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <script> [gpus] — retries while the pod answers "busy" (exit 3); log in gpurun_out/<script>.stdout
s=$(basename $2 .sh)
for try in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ -n "$3" ]; then gpurun --gpus $3 --timeout $1 -- bash $2 > gpurun_out/$s.stdout 2>&1; else gpurun --timeout $1 -- bash $2 > gpurun_out/$s.stdout 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $try tries"; exit $rc; fi
  sleep 120
done
echo "gave up"; exit 3
