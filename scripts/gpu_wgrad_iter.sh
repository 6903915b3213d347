#!/bin/bash
# development: wgrad_band cluster reduction -- parity, per-class time, step A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "wgrad" 2>&1 | tail -4
for v in 1 2 4; do
  echo "== CLUSTER=$v"
  SMSUT_WGRAD_CLUSTER=$v timeout 300 python scripts/conv_classes.py 10 2>&1 | head -7
done
timeout 900 bash scripts/gpu_ab.sh SMSUT_WGRAD_CLUSTER 1 2 4
