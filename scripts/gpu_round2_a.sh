#!/bin/bash
# GPU run A of round 2: kernel + parity tests (one process per file), bench (ours + reference arm), smoke.
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_kernels_gpu.py tests/test_parity_layers_gpu.py tests/test_modules_gpu.py
echo "== tests rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "== bench rc $?"
tail -c 1500 gpurun_out/bench_r2a.err | tail -5
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_r2a.json'))
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e'], d['roofline']['frac'], d.get('torch_cuda_context'), d['cpu_baseline'])
except Exception as e: print('bench parse failed', e)
PY
