#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_layers_gpu.py -m gpu -q -x -k "deterministic or headline" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -q -x -k "consis" 2>&1 | tail -4
timeout 900 bash scripts/gpu_ab.sh SMSUT_DEFER_WGRAD 0 1
