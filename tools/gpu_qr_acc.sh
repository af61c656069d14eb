#!/bin/bash
mkdir -p gpurun_out
for v in LXB_QR_TWOLEVEL=1 LXB_QR_TWOLEVEL=0; do
  env $v python tools/qr_tc_accuracy.py 32768 1024 2>&1 | tail -2
  env $v python tools/qr_tc_accuracy.py 65536 2048 2>&1 | tail -2
done | tee gpurun_out/r02_qr_accuracy.log
timeout 900 python -m pytest tests/test_direct_gpu.py -q -m gpu -k "qr" -x 2>&1 | tail -3
timeout 600 python bench.py --workload qr262k --no-cpu-baseline > gpurun_out/r02_bench_qr.json 2> gpurun_out/r02_bench_qr.err
python -c "import json; d=json.load(open('gpurun_out/r02_bench_qr.json')); print('RESULT', d['ms_per_step'], d['roofline']['frac'], d['parity'])"
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "qr" 2>&1 | tail -5
