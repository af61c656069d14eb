bash tools/gpu_qr_launches.sh > gpurun_out/launch_summary.txt 2>&1
bash tools/gpu_qr_ncu.sh update128_ws
head -12 gpurun_out/launch_summary.txt
