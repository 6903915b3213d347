bash scripts/gpu_tests.sh tests/test_kernels_gpu.py tests/test_modules_gpu.py || exit 1
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_r1final_ref.json 2> gpurun_out/bench_r1final_ref.err || exit 1
python bench.py > gpurun_out/bench_r1final.json 2> gpurun_out/bench_r1final.err || exit 1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/conv_classes.py 10 > gpurun_out/conv_classes_r1final.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/bench_r1final.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['roofline']['frac'], d['step_roofline']['frac'], d['cpu_baseline']['value'])"
