"""Multi-GPU row-sharded GMRES (needs >= 2 CUDA devices; skipped otherwise): launches
tests/dist_gmres_check.py under torchrun, which checks the fused peer-memory kernel against the
oracle and the single-GPU kernel."""
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
