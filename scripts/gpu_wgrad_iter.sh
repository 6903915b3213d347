#!/bin/bash
# development: wgrad_band variants -- parity, per-class time, step A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "wgrad" 2>&1 | tail -4
for cfg in "0 1" "1 1" "1 0"; do
  set -- $cfg
  echo "== MULTI=$1 M64=$2"
  SMSUT_WGRAD_MULTI=$1 SMSUT_WGRAD_M64=$2 timeout 300 python scripts/conv_classes.py 10 2>&1 | head -8
done
timeout 900 bash scripts/gpu_ab.sh SMSUT_WGRAD_MULTI 0 1
