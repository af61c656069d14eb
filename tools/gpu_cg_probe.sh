#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cg_gpu.py -q -m gpu > gpurun_out/r02_pytest_cg.log 2>&1; tail -3 gpurun_out/r02_pytest_cg.log
for v in "LXB_CG_PREFETCH=0 LXB_CG_STAGGER_NS=0" "LXB_CG_PREFETCH=1 LXB_CG_STAGGER_NS=0" "LXB_CG_PREFETCH=0 LXB_CG_STAGGER_NS=1500" "LXB_CG_PREFETCH=1 LXB_CG_STAGGER_NS=1500" "LXB_CG_PREFETCH=1 LXB_CG_STAGGER_NS=3000" "LXB_CG_PREFETCH=1 LXB_CG_STAGGER_NS=800"; do
  env $v timeout 600 python bench.py --workload cg256 --no-cpu-baseline --steps 30 > gpurun_out/r02_bench_cg.json 2> gpurun_out/r02_bench_cg.err
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_cg.json')); print('RESULT $v', d['ms_per_step'], d['roofline']['read_once_frac'], d['parity']['max_rel_err'])"
done
