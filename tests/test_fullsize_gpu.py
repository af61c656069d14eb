"""Full-size parity of the large single-system BASELINE configs against the CPU oracle:
C4 GMRES(20) on 32768^2 fp32 and C5 LSMR / QR on 262144 x 4096 fp32 (the sizes bench.py times; the
default fp32 QR path runs its trailing update on the tcgen05 tensor cores with the 3xTF32 split).
Slow (about a minute of host work per test: generating 10^9 normals and running the NumPy / LAPACK
oracle on a 4.3 GB operator), but part of the normal `-m gpu` run."""
import numpy as np
import pytest
import torch

import oracle
from tests.helpers import assert_close_tol, dev, host, tol_for

pytestmark = pytest.mark.gpu


def _ops():
    from lineax_b200 import _ops

    return _ops


def _normals(rng, shape, scale):
    a = rng.standard_normal(shape, dtype=np.float32)
    a *= np.float32(scale)
    return a


def test_c4_gmres_full_size_vs_oracle():
    """BASELINE configs[3]: the reference's easy generator (benchmarks/solver_speeds.py:146-152) at
    n = 32768: x within 1e-5, identical RESULTS code, num_steps within +-2 of the oracle."""
    n = 32768
    rng = np.random.default_rng(4)
    a = _normals(rng, (n, n), 1.0 / n)
    a[np.arange(n), np.arange(n)] += np.float32(2.0)
    xt = rng.standard_normal(n, dtype=np.float32)
    b = a @ xt
    A = torch.as_tensor(a).cuda()
    x, res, steps = _ops().gmres(A, dev(b), None, None, 1e-6, 1e-6, 10 * n, 20, 20, 0)
    x, res, steps = host(x), int(res), int(steps)
    del A
    torch.cuda.empty_cache()
    xr, rr, st = oracle.gmres(a, b, 1e-6, 1e-6)
    assert res == rr == 0
    assert abs(steps - st["num_steps"]) <= 2, (steps, st["num_steps"])
    # A = 2 I + N(0,1)/n: singular values within 2 +- 2/sqrt(n) -> cond ~ 1.01, flat north_star tolerance
    assert_close_tol(x, xr, tol_for(np.float32, 1.02, solver_tol=1e-6))
    assert_close_tol(x, xt, 4 * tol_for(np.float32, 1.02, solver_tol=1e-6), "solution vs x_true")


@pytest.fixture(scope="module")
def c5_problem():
    """C5 generator: A = N(0,1)/sqrt(m) (262144 x 4096 fp32), b = A x* + 0.1 N(0,1)."""
    m, n = 262144, 4096
    rng = np.random.default_rng(5)
    a = _normals(rng, (m, n), 1.0 / np.sqrt(m))
    xt = rng.standard_normal(n, dtype=np.float32)
    b = a @ xt + np.float32(0.1) * rng.standard_normal(m, dtype=np.float32)
    return a, b


def _normal_eq_residual(a, x, b):
    """||A^T (A x - b)||_inf / (||A||_F ||A x - b||_2), float64 accumulation, blocked over rows."""
    g = np.zeros(a.shape[1])
    r2 = 0.0
    x64 = x.astype(np.float64)
    for lo in range(0, a.shape[0], 32768):
        blk = a[lo:lo + 32768].astype(np.float64)
        r = blk @ x64 - b[lo:lo + 32768]
        g += blk.T @ r
        r2 += float(r @ r)
    return float(np.abs(g).max() / (np.sqrt(a.shape[1]) * np.sqrt(r2)))  # ||A||_F ~ sqrt(n) for this generator


def test_c5_lsmr_full_size_vs_oracle(c5_problem):
    a, b = c5_problem
    m, n = a.shape
    A = torch.as_tensor(a).cuda()
    x, res, steps, st = _ops().lsmr(A, dev(b), None, 1e-6, 1e-6, 1e8, 10 * n, 0)
    x, res, steps, st = host(x), int(res), int(steps), host(st)
    del A
    torch.cuda.empty_cache()
    xr, rr, s = oracle.lsmr(a, b, 1e-6, 1e-6)
    assert res == rr == 0
    assert abs(steps - s["num_steps"]) <= 2, (steps, s["num_steps"])
    assert int(st[0]) == s["istop"]
    # kappa(A) = (1 + sqrt(n/m)) / (1 - sqrt(n/m)) = 1.29 for this generator
    kap = (1 + np.sqrt(n / m)) / (1 - np.sqrt(n / m))
    assert_close_tol(x, xr, tol_for(np.float32, kap * kap, solver_tol=1e-6))
    assert _normal_eq_residual(a, x, b) < 1e-5


def test_c5_qr_full_size_vs_lapack(c5_problem):
    a, b = c5_problem
    m, n = a.shape
    A = torch.as_tensor(a).cuda()
    aq, taus = _ops().qr_factor(A[None])
    x = host(_ops().qr_solve(aq, taus, dev(b[None]), False))[0]
    r_gpu = np.triu(host(aq[0, :n]))
    taus_gpu = host(taus)[0]
    del A, aq
    torch.cuda.empty_cache()
    st = oracle.qr_init(a)  # LAPACK sgeqrf: what jaxlib's CPU backend calls
    (a_ref, taus_ref), _ = st
    xr = oracle.qr_compute(st, b)
    r_ref = np.triu(a_ref[:n])
    del a_ref, st
    # least-squares solution: flat north_star tolerance (kappa^2 = 1.66 for this generator)
    kap = (1 + np.sqrt(n / m)) / (1 - np.sqrt(n / m))
    tol = tol_for(np.float32, kap * kap)
    assert tol == 1e-5
    assert_close_tol(x, xr, tol)
    assert _normal_eq_residual(a, x, b) < 1e-5
    # R and taus follow LAPACK's sign conventions; both factorisations carry the rounding of 262144-term
    # column norms, so R is compared against float64 truth as well: R^T R = A^T A on the leading 64 columns
    assert np.max(np.abs(r_gpu - r_ref)) <= 2e-5 * np.max(np.abs(r_ref)), np.max(np.abs(r_gpu - r_ref))
    assert np.max(np.abs(taus_gpu - taus_ref)) <= 2e-5
    a64 = a[:, :64].astype(np.float64)
    gram = a64.T @ a64
    r64 = r_gpu[:64, :64].astype(np.float64)
    assert np.max(np.abs(r64.T @ r64 - gram)) <= 1e-5 * np.max(np.abs(gram))
