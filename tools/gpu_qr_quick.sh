#!/bin/bash
# quick check of the large-QR path: accuracy vs float64 on two shapes and the qr262k bench line
mkdir -p gpurun_out
{
  timeout 120 python tools/qr_tc_accuracy.py 32768 1024
  timeout 120 python tools/qr_tc_accuracy.py 20000 1100
  timeout 200 python bench.py --workload qr262k --no-cpu-baseline --steps 5 --warmup 3
} > gpurun_out/qr_quick.log 2>&1
grep -v '^{' gpurun_out/qr_quick.log | tail; grep -o '"ms_per_step": [0-9.]*' gpurun_out/qr_quick.log
