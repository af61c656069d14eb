"""Multi-GPU row-sharded GMRES and LSMR (need >= 2 CUDA devices; skipped otherwise): launch
tests/dist_gmres_check.py / tests/dist_lsmr_check.py under torchrun, which check the fused
peer-memory kernels against the oracle and the single-GPU kernels."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_row_sharded_gmres_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gmres_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "DIST_GMRES_ALL_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_row_sharded_lsmr_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "dist_lsmr_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "DIST_LSMR_ALL_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
