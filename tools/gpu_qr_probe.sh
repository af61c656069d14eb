#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_direct_gpu.py -q -m gpu -k "qr" -x > gpurun_out/r02_pytest_qr.log 2>&1; tail -15 gpurun_out/r02_pytest_qr.log
for v in LXB_QR_TWOLEVEL=1 LXB_QR_TWOLEVEL=0; do
  env $v timeout 600 python bench.py --workload qr262k --no-cpu-baseline > gpurun_out/r02_bench_qr_$v.json 2> gpurun_out/r02_bench_qr_$v.err
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_qr_$v.json')); print('RESULT $v', d['ms_per_step'], d['roofline']['frac'], d['parity'])" || tail -5 gpurun_out/r02_bench_qr_$v.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_qr262k.csv python bench.py --workload qr262k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_qr.log 2>&1
