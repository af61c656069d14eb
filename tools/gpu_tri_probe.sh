#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_direct_gpu.py tests/test_api_gpu.py -q -m gpu -k "tridiag or Tridiag" > gpurun_out/r02_pytest_tri.log 2>&1; tail -5 gpurun_out/r02_pytest_tri.log
timeout 600 python bench.py --workload tridiag512 --no-cpu-baseline > gpurun_out/r02_bench_tri_default.json 2> gpurun_out/r02_bench_tri_default.err
python -c "import json; d=json.load(open('gpurun_out/r02_bench_tri_default.json')); print('RESULT', d['ms_per_step'], d['roofline']['frac'], d['parity'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tridiag_warp -s 3 -c 1 -o gpurun_out/r02_prof_tridiag -f python bench.py --workload tridiag512 --steps 3 --no-cpu-baseline > gpurun_out/r02_ncu_tri.log 2>&1; tail -2 gpurun_out/r02_ncu_tri.log
