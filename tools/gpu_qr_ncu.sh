#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wbig_tc_kernel -s 1 -c 1 -o gpurun_out/r02_prof_qr_wbig -f python bench.py --workload qr262k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_qr1.log 2>&1
tail -3 gpurun_out/r02_ncu_qr1.log
