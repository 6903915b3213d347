#!/bin/bash
# development: conv_tc weight multicast -- parity, per-class time, step A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "multicast or conv_tc_fprop" 2>&1 | tail -6
for v in 0 2 4; do
  echo "== TC_MCAST=$v"
  SMSUT_TC_MCAST=$v timeout 300 python scripts/conv_classes.py 10 2>&1 | tail -9
done
timeout 900 bash scripts/gpu_ab.sh SMSUT_TC_MCAST 0 2 4
