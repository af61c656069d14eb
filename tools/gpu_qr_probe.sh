#!/bin/bash
# QR two-level path: accuracy vs float64, the QR parity tests, and the qr262k bench line.
mkdir -p gpurun_out
{
  timeout 120 python tools/qr_tc_accuracy.py 32768 1024
  timeout 120 python tools/qr_tc_accuracy.py 20000 1100
  timeout 200 python bench.py --workload qr262k --no-cpu-baseline --steps 5 --warmup 3
  timeout 300 python -m pytest tests/test_direct_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "qr or QR" 2>&1 | tail -5
} > gpurun_out/qr_probe.log 2>&1
grep -v '^{' gpurun_out/qr_probe.log | tail; grep -o '"ms_per_step": [0-9.]*' gpurun_out/qr_probe.log
