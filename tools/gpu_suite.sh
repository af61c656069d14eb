#!/bin/bash
# Full GPU check: pytest -m gpu, smoke, default bench (headline + extras). Run under gpurun.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -q -m gpu --maxfail=60 --durations=15 ${PYTEST_ARGS} > gpurun_out/r02_pytest_gpu.log 2>&1; tail -40 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 1200 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 600 gpurun_out/r02_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print('HEAD', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline',{}) and d['cpu_baseline']['value'], d['parity'])
for k,v in d.get('extras',{}).items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['e2e']['value'], v['cpu_baseline'] and v['cpu_baseline']['value'], v['parity'], v['wall_s'])
PY
