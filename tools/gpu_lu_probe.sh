#!/bin/bash
# GPU probe for the TMA LU kernels: parity + timing of the variants (+ optional ncu capture). Run under gpurun.
# usage: bash tools/gpu_lu_probe.sh "<VAR=val|default> ..." [ncu-variant]
set -x
mkdir -p gpurun_out
for v in $1; do
  tag=$(echo $v | tr '=' '_')
  if [ "$v" = default ]; then env_=""; else env_="$v"; fi
  env $env_ timeout 600 python -m pytest tests/test_lu_gpu.py -x -q > gpurun_out/r02_pytest_lu_$tag.log 2>&1; tail -3 gpurun_out/r02_pytest_lu_$tag.log
  env $env_ timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_lu_$tag.json 2> gpurun_out/r02_bench_lu_$tag.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r02_bench_lu_$tag.json')); print('RESULT $tag', d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done
if [ -n "$2" ]; then
  tag=$(echo $2 | tr '=' '_')
  if [ "$2" = default ]; then env_=""; else env_="$2"; fi
  env $env_ timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu32_ -s 5 -c 2 -o gpurun_out/r02_prof_lu32_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_lu.log 2>&1
  tail -3 gpurun_out/r02_ncu_lu.log
fi
