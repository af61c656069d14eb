#!/bin/bash
# usage: bash tools/gpu_multi_quick.sh N   -- sharded-kernel parity + gmres/lsmr/qr timings on N GPUs
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x > gpurun_out/r02_pytest_multigpu_$N.log 2>&1; tail -8 gpurun_out/r02_pytest_multigpu_$N.log
for w in gmres32k lsmr262k qr262k; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r02_scale_${w}_$N.json 2> gpurun_out/r02_scale_${w}_$N.err
  python - <<PY
import json
lines=[l for l in open('gpurun_out/r02_scale_${w}_$N.json') if l.startswith('{')]
d=json.loads(lines[-1]); print('RESULT', '$w', $N, d['ms_per_step'], d['roofline']['frac'], d['config']['num_steps'], d['parity'].get('full_size'))
PY
done
