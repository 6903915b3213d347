#!/bin/bash
# multi-GPU bench lines: usage gpu_multi.sh N [extra workloads...]
N=$1; shift
mkdir -p gpurun_out
run() {  # workload tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 20 --warmup 3 $1 > gpurun_out/bench_r2_$2${N}gpu.json 2> gpurun_out/bench_$2${N}gpu.err
  echo "== N=$N $2 rc $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2_$2${N}gpu.json'))
    print(round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), 'scaling', d['scaling'], 'clocks', d.get('clocks'))
    if 'parity' in d: print('parity', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['parity'].items() if k!='losses_rel'}, 'max loss rel', max(d['parity']['losses_rel'].values()))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_$2${N}gpu.err').read()[-3000:])
PY
}
run "" ""
for w in "$@"; do run "--workload $w" "${w}_"; done
