#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x > gpurun_out/r02_pytest_multigpu.log 2>&1; tail -12 gpurun_out/r02_pytest_multigpu.log
