#!/bin/bash
# usage: bash tools/gpu_multi_perf.sh N "workloads"  -- timings only
set -x
N=${1:-2}
mkdir -p gpurun_out
for w in $2; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r02_scale_${w}_$N.json 2> gpurun_out/r02_scale_${w}_$N.err
  python - <<PY
import json
lines=[l for l in open('gpurun_out/r02_scale_${w}_$N.json') if l.startswith('{')]
d=json.loads(lines[-1]); print('RESULT', '$w', $N, d['ms_per_step'], d['roofline']['frac'], d['config']['num_steps'], d['parity'].get('full_size'))
PY
done
