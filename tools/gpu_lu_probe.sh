#!/bin/bash
# GPU probe for the TMA LU kernel: parity, timing of the variants, one ncu capture. Run under gpurun.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_lu_gpu.py -x -q > gpurun_out/r02_pytest_lu.log 2>&1; tail -5 gpurun_out/r02_pytest_lu.log
for v in default LXB_LU_MINB LXB_LU_NONUNI; do
  if [ "$v" = default ]; then env_=""; else env_="$v=3"; fi
  env $env_ timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_lu_$v.json 2> gpurun_out/r02_bench_lu_$v.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r02_bench_lu_$v.json')); print('$v', d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu32_tma -s 5 -c 2 -o gpurun_out/r02_prof_lu32tma -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_lu.log 2>&1
tail -3 gpurun_out/r02_ncu_lu.log
