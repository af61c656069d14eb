"""Grid-cooperative tier (one large system at a time on all SMs): parity vs the oracle for the
BASELINE large-system configs at reduced size, plus agreement with the CTA tier."""
import numpy as np
import pytest

import oracle
from oracle import gen
from tests.helpers import assert_close, assert_close_tol, dev, host, max_cond, tol_for
from tests.test_cg_gpu import run_cg
from tests.test_krylov_gpu import run_bicgstab, run_gmres, run_lsmr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,dtype,tol", [(512, np.float32, 1e-6), (1024, np.float64, 1e-12), (2050, np.float32, 1e-6)])
def test_cg_single_large(n, dtype, tol):
    a, b, xt = gen.easy_problem(n, n, dtype, spd=True)
    x, res, steps = run_cg(a[None], b[None], tol, tol)
    xr, rr, st = oracle.cg(a, b, tol, tol)
    assert res[0] == rr == 0 and abs(int(steps[0]) - st["num_steps"]) <= 2
    assert_close_tol(x[0], xr, tol_for(dtype, max_cond(a), solver_tol=tol))


def test_cg_grid_variants():
    a, b, _ = gen.easy_problem(9, 600, np.float64, spd=True, batch=3)
    M = np.stack([np.diag(1.0 / np.diag(a[i])) for i in range(3)])
    y0 = np.random.default_rng(0).standard_normal((3, 600))
    x, res, steps = run_cg(-a, b, 1e-10, 1e-10, nsd=True, precond=M, y0=y0, stabilise_every=3)
    for i in range(3):
        xr, rr, st = oracle.cg(-a[i], b[i], 1e-10, 1e-10, is_nsd=True, preconditioner=M[i], y0=y0[i],
                               stabilise_every=3)
        assert res[i] == rr and abs(int(steps[i]) - st["num_steps"]) <= 2
        assert_close_tol(x[i], xr, tol_for(np.float64, max_cond(a[i]), solver_tol=1e-10))
    x, res, steps = run_cg(a[:1], b[:1], 0.0, 0.0, max_steps=4)
    xr, rr, st = oracle.cg(a[0], b[0], 0.0, 0.0, max_steps=4)
    assert res[0] == rr == 0 and steps[0] == 4
    assert_close_tol(x[0], xr, tol_for(np.float64, max_cond(a[0])))  # exactly 4 steps on both sides


@pytest.mark.parametrize("n,dtype,tol", [(512, np.float32, 1e-6), (1500, np.float64, 1e-12)])
def test_bicgstab_single_large(n, dtype, tol):
    a, b, _ = gen.easy_problem(n + 1, n, dtype, spd=False)
    x, res, steps = run_bicgstab(a[None], b[None], tol, tol)
    xr, rr, st = oracle.bicgstab(a, b, tol, tol)
    assert abs(int(steps[0]) - st["num_steps"]) <= 2
    if steps[0] == st["num_steps"]:
        assert res[0] == rr
    assert_close_tol(x[0], xr, tol_for(dtype, max_cond(a), solver_tol=tol))


@pytest.mark.parametrize("n,dtype,tol", [(512, np.float32, 1e-6), (2048, np.float32, 1e-6), (1000, np.float64, 1e-12)])
def test_gmres_single_large(n, dtype, tol):
    """C4 generator (easy nonsymmetric) at reduced n: 3-4 restarts."""
    a, b, _ = gen.easy_problem(n + 2, n, dtype, spd=False)
    x, res, steps = run_gmres(a[None], b[None], tol, tol)
    xr, rr, st = oracle.gmres(a, b, tol, tol)
    assert res[0] == rr == 0 and abs(int(steps[0]) - st["num_steps"]) <= 2
    assert st["num_steps"] in (3, 4, 5)
    assert_close_tol(x[0], xr, tol_for(dtype, max_cond(a), solver_tol=tol))


def test_gmres_grid_failure_codes_and_precond():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((600, 600))
    bb = a @ rng.standard_normal(600)
    x, res, steps = run_gmres(a[None], bb[None], 1e-10, 1e-10, restart=2)
    xr, rr, st = oracle.gmres(a, bb, 1e-10, 1e-10, restart=2)
    assert rr != 0 and res[0] == rr and abs(int(steps[0]) - st["num_steps"]) <= 2
    a2, b2, _ = gen.easy_problem(5, 700, np.float64, spd=False)
    M = np.diag(1.0 / np.diag(a2))
    x, res, steps = run_gmres(a2[None], b2[None], 1e-10, 1e-10, precond=M[None], restart=10)
    xr, rr, st = oracle.gmres(a2, b2, 1e-10, 1e-10, preconditioner=M, restart=10)
    assert res[0] == rr == 0 and abs(int(steps[0]) - st["num_steps"]) <= 2
    assert_close_tol(x[0], xr, tol_for(np.float64, max_cond(a2), solver_tol=1e-10))


@pytest.mark.parametrize("shape,dtype,tol", [((16384, 256), np.float32, 1e-6), ((4096, 512), np.float64, 1e-12),
                                             ((300, 3000), np.float32, 1e-6), ((1030, 1030), np.float64, 1e-10)])
def test_lsmr_single_large(shape, dtype, tol):
    m, n = shape
    a, b, _ = gen.tall_lstsq(m + n, m, n, dtype) if m >= n else (None, None, None)
    if m == n:  # well-conditioned square system (reference's easy generator): a handful of steps
        a, b, _ = gen.easy_problem(m, n, dtype, spd=False)
    elif a is None:
        rng = np.random.default_rng(m)
        a = (rng.standard_normal((m, n)) / np.sqrt(n)).astype(dtype)
        b = rng.standard_normal(m).astype(dtype)
    x, res, steps, st = run_lsmr(a[None], b[None], tol, tol)
    xr, rr, s = oracle.lsmr(a, b, tol, tol)
    assert res[0] == rr and abs(int(steps[0]) - s["num_steps"]) <= 2
    xl = np.linalg.lstsq(a.astype(np.float64), b.astype(np.float64), rcond=None)[0]
    kap = max_cond(a)
    t = tol_for(dtype, kap * kap if m != n else kap, solver_tol=tol)
    assert_close_tol(x[0], xl, t, "solution vs float64 lstsq")
    assert_close_tol(x[0], xr, t)
