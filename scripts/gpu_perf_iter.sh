#!/bin/bash
# perf iteration on one B200: the kernel parity tests that cover what changed, the determinism test, a quick bench line
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -4
python -m pytest tests/test_parity_layers_gpu.py -m gpu -q -x -k "deterministic or unet_layers or discriminator" 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-context > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_iter.json'))
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, 'e2e', d['e2e']['value'], 'conv frac', d['roofline']['frac'])
    for r in d['roofline']['per_class']: print("  %4d->%3d @%3d %6.1f us %6.1f TF/s %6.0f GB/s ideal %5.1f"%(r['cin'],r['cout'],r['hw'],r['us'],r['tflops'],r['gbs'],r['ideal_us']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_iter.err').read()[-2000:])
PY
