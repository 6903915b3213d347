#!/bin/bash
# Run the GPU test files one process each (a trapped kernel poisons only its own CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -m gpu -q --timeout=300 -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  r=$?
  echo "== $f -> exit $r"
  tail -n 25 "gpurun_out/$name.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
