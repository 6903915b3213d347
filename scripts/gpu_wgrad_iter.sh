#!/bin/bash
# development: wgrad_band CTAs per SM -- per-class time, step A/B
mkdir -p gpurun_out
for v in 1 2; do
  echo "== PER_SM=$v"
  SMSUT_WGRAD_BAND_PER_SM=$v timeout 300 python scripts/conv_classes.py 10 2>&1 | head -7
done
timeout 900 bash scripts/gpu_ab.sh SMSUT_WGRAD_BAND_PER_SM 1 2
