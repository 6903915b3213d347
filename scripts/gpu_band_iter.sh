#!/bin/bash
# development: conv_band CTAs per SM / ring depth -- per-class time, step A/B
mkdir -p gpurun_out
for cfg in "3 0" "2 0" "2 10" "1 0" "1 12"; do
  set -- $cfg
  echo "== BAND_PER_SM=$1 SLOTS=$2"
  if [ "$2" = "0" ]; then unset SMSUT_BAND_SLOTS; else export SMSUT_BAND_SLOTS=$2; fi
  SMSUT_BAND_PER_SM=$1 timeout 300 python scripts/conv_classes.py 10 2>&1 | head -6
  SMSUT_BAND_PER_SM=$1 python bench.py --steps 30 --warmup 3 --no-context --no-check > gpurun_out/abb.json 2> gpurun_out/abb.err
  python -c "
import json
d=json.load(open('gpurun_out/abb.json')); print('  step', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'slices/s')"
done
