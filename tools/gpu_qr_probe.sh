#!/bin/bash
set -x
mkdir -p gpurun_out
make -C lineax_b200/csrc -j16 EXTRA="-DLXB_QR_WTC_EXPERIMENT" -B > gpurun_out/r02_qr_build.log 2>&1; tail -2 gpurun_out/r02_qr_build.log
LXB_QR_WTC=1 timeout 900 python -m pytest tests/test_direct_gpu.py -q -m gpu -k "qr" > gpurun_out/r02_pytest_qr_wtc.log 2>&1; tail -12 gpurun_out/r02_pytest_qr_wtc.log
for v in LXB_QR_WTC=1 LXB_QR_WTC=0; do
  env $v timeout 600 python bench.py --workload qr262k --no-cpu-baseline > gpurun_out/r02_bench_qr_$v.json 2> gpurun_out/r02_bench_qr_$v.err
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_qr_$v.json')); print('RESULT $v', d['ms_per_step'], d['roofline']['frac'], d['parity'])"
done
