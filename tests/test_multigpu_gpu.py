"""Row-sharded GMRES and LSMR.  The fused peer-memory kernels are exercised on EVERY box: as a 1-rank group
(symmetric-memory rendezvous with itself; every exchange round, flag and fold of the kernels still runs) and,
where >= 2 CUDA devices exist, across 2 GPUs.  Launch
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


def _run(script, nproc, port, token):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert token in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


def test_row_sharded_gmres_one_rank_group():
    """gmres_dist_kernel on a 1-rank group (runs on the driver's single-GPU box)."""
    _run("dist_gmres_check.py", 1, 29535, "DIST_GMRES_ALL_OK")


def test_row_sharded_lsmr_one_rank_group():
    """lsmr_dist_kernel on a 1-rank group (runs on the driver's single-GPU box)."""
    _run("dist_lsmr_check.py", 1, 29536, "DIST_LSMR_ALL_OK")


def test_row_sharded_qr_and_linear_solve_entry_one_rank_group():
    """TSQR + the `lx.linear_solve(RowShardedMatrixLinearOperator, ...)` entry points on a 1-rank group."""
    _run("dist_qr_check.py", 1, 29537, "DIST_QR_ALL_OK")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_row_sharded_qr_two_gpus():
    _run("dist_qr_check.py", 2, 29538, "DIST_QR_ALL_OK")


def test_tsqr_combine_math_single_gpu():
    """The TSQR reduction itself on ONE GPU: factor 4 row blocks separately, stack their R factors and
    reduced right-hand sides, solve the stacked system -- must equal the least-squares solution."""
    import numpy as np

    from lineax_b200 import _ops
    from oracle import gen

    m, n, parts = 16384, 256, 4
    a, b, _ = gen.tall_lstsq(21, m, n, np.float32)
    A, B = torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda()
    rs, cs = [], []
    for p in range(parts):
        lo, hi = p * m // parts, (p + 1) * m // parts
        aq, taus = _ops.qr_factor(A[lo:hi])
        cs.append(_ops.qr_apply_qt(aq, taus, B[lo:hi]))
        rs.append(torch.triu(aq[:n]))
    aq2, taus2 = _ops.qr_factor(torch.cat(rs))
    x = _ops.qr_solve(aq2, taus2, torch.cat(cs), False).cpu().numpy()
    xl = np.linalg.lstsq(a.astype(np.float64), b.astype(np.float64), rcond=None)[0]
    assert np.abs(x - xl).max() / np.abs(xl).max() < 1e-5
