#!/bin/bash
# Multi-GPU check: sharded-kernel parity tests + the default bench (headline + extras) under torchrun.
# usage: bash tools/gpu_multi.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -q -m gpu > gpurun_out/r02_pytest_multigpu_$N.log 2>&1; tail -8 gpurun_out/r02_pytest_multigpu_$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_gpus$N.json 2> gpurun_out/r02_bench_gpus$N.err; tail -c 1500 gpurun_out/r02_bench_gpus$N.err
python - <<PY
import json
lines=[l for l in open('gpurun_out/r02_bench_gpus$N.json') if l.startswith('{')]
d=json.loads(lines[-1])
print('HEAD', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
for k,v in d.get('extras',{}).items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    print(k, v['scaling'], v['value'], v['ms_per_step'], v['roofline']['frac'], v['e2e']['value'], v['parity'].get('full_size'), v['wall_s'])
PY
