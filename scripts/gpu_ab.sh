#!/bin/bash
# A/B of an environment knob: usage gpu_ab.sh KNOB v1 v2 ...   (bench.py --no-context, 30 steps each)
knob=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $knob=$v python bench.py --steps 30 --warmup 3 --no-context > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_$v.json')); print("$knob=$v", round(d['value'],1), 'slices/s', round(d['ms_per_step'],3), 'ms', d['launches_per_step'], 'launches; e2e', round(d['e2e']['value'],1))
except Exception as e: print("$knob=$v failed", e)
PY
done
