#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qr_wbig_tc_kernel -s 8 -c 1 -o gpurun_out/r02_prof_qr_wbig -f python bench.py --workload qr262k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_qr1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qr_update128 -s 4 -c 1 -o gpurun_out/r02_prof_qr_up128 -f python bench.py --workload qr262k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_qr2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
