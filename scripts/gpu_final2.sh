#!/bin/bash
# per-SM rule of wgrad_band (env=1: one CTA per SM everywhere; unset: two for the single-issuer 1x1 layers), the ncu
# capture of the conv classes with the weight multicast on (evidence for the measured negative), wgrad tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "wgrad" 2>&1 | tail -3
SMSUT_WGRAD_BAND_PER_SM=1 python bench.py --steps 30 --warmup 3 --no-context > gpurun_out/ab_p1.json 2> gpurun_out/ab_p1.err
python bench.py --steps 30 --warmup 3 --no-context > gpurun_out/ab_prule.json 2> gpurun_out/ab_prule.err
python - <<PY
import json
for k in ('p1','prule'):
    d=json.load(open(f'gpurun_out/ab_{k}.json')); print(k, round(d['value'],1), 'slices/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1))
PY
SMSUT_TC_MCAST=4 python scripts/conv_ncu.py > /dev/null 2>&1 || exit 1
SMSUT_TC_MCAST=4 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel' -o gpurun_out/conv_mcast -f \
    python scripts/conv_ncu.py > gpurun_out/ncu_conv_mcast.log 2>&1
ncu -i gpurun_out/conv_mcast.ncu-rep --page raw --csv > gpurun_out/conv_mcast_raw.csv 2>/dev/null
rm -f gpurun_out/conv_mcast.ncu-rep
tail -2 gpurun_out/ncu_conv_mcast.log
