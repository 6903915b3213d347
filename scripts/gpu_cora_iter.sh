#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "coranet" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -q -x -k "coranet" 2>&1 | tail -15
for v in 64 32; do
  echo "== TC_MAXCC=$v"
  SMSUT_TC_MAXCC=$v timeout 300 python scripts/conv_classes.py 10 2>&1 | tail -9
done
timeout 900 bash scripts/gpu_ab.sh SMSUT_TC_MAXCC 64 32
