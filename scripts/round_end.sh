#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): GPU tests, both bench arms, then -- only after those exited 0
# without a profiler -- the ncu launch list of one eager iteration and a --set full capture of the IN backward pair.
tag=${1:-r1}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_kernels_gpu.py tests/test_modules_gpu.py || exit 1
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err || exit 1
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/profile_step.py > gpurun_out/profile_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/launches_${tag}.csv \
    python scripts/profile_step.py > gpurun_out/ncu_run.log 2>&1
python scripts/in_probe2.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:in_bwd -o gpurun_out/inbwd_${tag} -f \
    python scripts/in_probe2.py > gpurun_out/ncu_inbwd.log 2>&1
ncu -i gpurun_out/inbwd_${tag}.ncu-rep --page raw --csv > gpurun_out/inbwd_${tag}_raw.csv 2>/dev/null
tail -c 400 gpurun_out/bench_${tag}.json | head -c 200; echo
