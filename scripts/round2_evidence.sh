#!/bin/bash
# Round-2 evidence on one B200 (through gpurun): GPU tests, both bench arms and the two extra workloads -- and only after
# those exited 0 without a profiler -- the ncu launch list (time + DRAM bytes) of one eager iteration, the --set full
# capture of the conv kernels on the 13 conv classes, per-class microseconds and the kernel timeline of the graph replay.
tag=${1:-r2}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_kernels_gpu.py tests/test_input_pipeline.py tests/test_parity_layers_gpu.py tests/test_modules_gpu.py
echo "== tests rc $?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${tag}_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc $?"
python bench.py --workload mean_teacher_512 --steps 20 --warmup 3 > gpurun_out/bench_${tag}_mt512_1gpu.json 2> gpurun_out/mt.err; echo "mt rc $?"
python bench.py --workload unet_infer --steps 10 > gpurun_out/bench_${tag}_unet_infer.json 2> gpurun_out/ui.err; echo "infer rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python scripts/conv_classes.py 10 > gpurun_out/${tag}_conv_classes_us.txt 2>&1; tail -3 gpurun_out/${tag}_conv_classes_us.txt
python scripts/timeline.py > gpurun_out/${tag}_timeline.txt 2>&1; tail -3 gpurun_out/${tag}_timeline.txt
python scripts/profile_step.py > gpurun_out/profile_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3400 --csv \
    --log-file gpurun_out/launches_${tag}.csv python scripts/profile_step.py > gpurun_out/ncu_run.log 2>&1
python scripts/conv_ncu.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'conv_band_kernel|conv_tc_kernel' -o gpurun_out/conv_${tag} -f \
    python scripts/conv_ncu.py > gpurun_out/ncu_conv.log 2>&1
ncu -i gpurun_out/conv_${tag}.ncu-rep --page raw --csv > gpurun_out/conv_${tag}_raw.csv 2>/dev/null
rm -f gpurun_out/conv_${tag}.ncu-rep
python scripts/wgrad_ncu.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'wgrad_band_kernel' -o gpurun_out/wgrad_${tag} -f \
    python scripts/wgrad_ncu.py > gpurun_out/ncu_wgrad.log 2>&1
ncu -i gpurun_out/wgrad_${tag}.ncu-rep --page raw --csv > gpurun_out/wgrad_${tag}_raw.csv 2>/dev/null
rm -f gpurun_out/wgrad_${tag}.ncu-rep
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${tag}_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, 'e2e', d['e2e']['value'], 'conv frac', d['roofline']['frac'], d.get('torch_cuda_context'))
r=json.load(open('gpurun_out/bench_${tag}_reference_arm.json')); print('reference arm', r['value'], r['steps'], r['cpu_baseline']['cores'])
PY
