#!/bin/bash
# one full ncu capture of a QR trailing-update kernel (KERNEL regex as $1, default the warp-specialised update)
K=${1:-update128_ws}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/r02_prof_qr_$K -f python bench.py --workload qr262k --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_qr1.log 2>&1
tail -3 gpurun_out/r02_ncu_qr1.log
