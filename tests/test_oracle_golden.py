"""Pins the CPU oracle against every closed-form / hard-coded vector the reference's own tests
hold for the hot path and against LAPACK / numpy.linalg as the reference's
test-suite does everywhere else (tests/test_well_posed.py:31-52).  Runs without a GPU."""
import numpy as np
import pytest

import oracle
from oracle import RESULTS, clib, gen


def rand_cond(rng, n, cond=1000.0, spd=False):
    while True:
        m = rng.standard_normal((n, n))
        if spd:
            m = m @ m.T
        if np.linalg.cond(m) < cond:
            return m


def test_kat_pytree_2x2():
    """tests/test_solve.py:41-48: [[1, 5], [-2, -2]] x = [3, 4] -> [-3.25, 1.25]."""
    a = np.array([[1.0, 5.0], [-2.0, -2.0]])
    b = np.array([3.0, 4.0])
    for x in (oracle.lu_compute(oracle.lu_init(a), b), oracle.qr_compute(oracle.qr_init(a), b),
              clib.lu_factor_solve(a[None], b[None])[0][0]):
        assert np.allclose(x, [-3.25, 1.25], atol=1e-14)


def test_kat_diagonal_pytree():
    """tests/test_solve.py:51-61."""
    d = np.array([8.0, 1, 2, 3, 4, 5, 6])
    y = np.array([4.0, 7, 8, 9, 2, 10, 12])
    assert np.allclose(oracle.diagonal_compute(d, y, well_posed=True), [0.5, 7, 4, 3, 0.5, 2, 2])
    assert np.allclose(oracle.diagonal_compute(d, y), [0.5, 7, 4, 3, 0.5, 2, 2])


def test_kat_triangular_mixed():
    """tests/test_solve.py:103-112: lower-triangular [[1, 0], [-2, -2]] x = [3, 4] -> [3, -5]."""
    a = np.array([[1.0, 0.0], [-2.0, -2.0]])
    assert np.allclose(oracle.triangular_compute(a, np.array([3.0, 4.0]), lower=True), [3.0, -5.0])


def test_kat_gmres_restart2_fails():
    """tests/test_singular.py:75-99: the hard-coded 4x4 with restart=2 must not be `successful`."""
    matrix = np.array([
        [0.15892892, 0.05884365, -0.60427412, 0.1891916],
        [-1.5484863, 0.93608822, 1.94888868, 1.37069667],
        [0.62687318, -0.13996738, -0.6824359, 0.30975754],
        [-0.67428635, 1.52372255, -0.88277754, 0.69633816],
    ])
    true_x = np.array([0.51383273, 1.72983427, -0.43251078, -1.11764668])
    x, res, st = oracle.gmres(matrix, matrix @ true_x, 1e-10, 1e-10, restart=2)
    assert res != RESULTS.successful
    # tests/test_singular.py:56-72: 100x100 Gaussian, restart=2 -> failure reported
    rng = np.random.default_rng(0)
    a = rng.standard_normal((100, 100))
    x, res, st = oracle.gmres(a, a @ rng.standard_normal(100), 1e-10, 1e-10, restart=2)
    assert res != RESULTS.successful


def test_kat_gmres_large_restart():
    """tests/test_solve.py:24-38: restart=100 on a 100x100 Gaussian converges to true_x."""
    rng = np.random.default_rng(1)
    a = rng.standard_normal((100, 100))
    xt = rng.standard_normal(100)
    x, res, st = oracle.gmres(a, a @ xt, 1e-10, 1e-10, restart=100)
    assert res == RESULTS.successful and np.allclose(x, xt, atol=1e-6)


def test_kat_lsmr_diag():
    """tests/test_lsmr.py:7-30."""
    ill = np.diag([1e8, 1e6, 1e4, 1e2, 1.0])
    well = np.diag([2.0, 4.0, 5.0, 8.0, 10.0])
    sing = np.diag([0.0, 4.0, 5.0, 8.0, 10.0])
    for m in (ill, well, sing):
        x, res, st = oracle.lsmr(m, np.zeros(5), 1e-10, 1e-10)
        assert np.all(x == 0) and res == RESULTS.successful
    e0 = np.zeros(5)
    e0[0] = 1
    x, res, st = oracle.lsmr(sing, e0, 1e-10, 1e-10)
    assert np.all(x == 0)
    x, res, st = oracle.lsmr(ill, np.ones(5), 1e-10, 1e-10, conlim=1e3)
    assert res == RESULTS.conlim and st["istop"] == 3


@pytest.mark.parametrize("solver", ["cg", "bicgstab", "gmres", "lsmr"])
def test_kat_max_steps_only_poisson(solver):
    """tests/test_solve.py:174-194: rtol=atol=0, max_steps=2 on Poisson(100): no failure."""
    p = gen.poisson_matrix(100)
    rhs = np.random.default_rng(0).standard_normal(100)
    fn = getattr(oracle, solver)
    kw = {"is_nsd": True} if solver == "cg" else {}
    x, res, st = fn(p, rhs, 0.0, 0.0, max_steps=2, **kw)
    assert res == RESULTS.successful and st["num_steps"] == 2


@pytest.mark.parametrize("solver,steps", [("gmres", 4), ("bicgstab", 3), ("cg", 3)])
def test_kat_exact_preconditioner(solver, steps):
    """tests/test_adjoint.py:84-130: uniform(10x10) + diag(i^6) with the exact inverse as
    preconditioner converges within max_steps (no throw)."""
    rng = np.random.default_rng(123)
    A = rng.uniform(size=(10, 10)) + np.diag(np.arange(10.0) ** 6)
    b = rng.uniform(size=10)
    kw = {}
    if solver == "cg":
        A = A.T @ A
    if solver == "gmres":
        kw["restart"] = 1
    x, res, st = getattr(oracle, solver)(A, b, 1e-12, 1e-12, max_steps=steps,
                                         preconditioner=np.linalg.inv(A), **kw)
    assert res == RESULTS.successful
    assert np.allclose(A @ x, b, atol=1e-8)


def test_nonfinite_postprocess():
    """tests/test_solve.py:249-261 and lineax/_solve.py:104-123."""
    for vec in ([1.0, np.inf], [1.0, np.nan], [np.nan, np.inf]):
        b = np.array(vec)
        x = oracle.diagonal_compute(np.ones(2), b, well_posed=True)
        assert oracle.postprocess(x, RESULTS.successful, b) == RESULTS.nonfinite_input
    assert oracle.postprocess(np.array([np.nan]), RESULTS.successful, np.array([1.0])) == RESULTS.singular
    assert oracle.postprocess(np.array([1.0]), RESULTS.max_steps_reached, np.array([np.nan])) == RESULTS.max_steps_reached


@pytest.mark.parametrize("name", ["lu", "qr", "cholesky", "cg", "bicgstab", "gmres", "lsmr", "tridiagonal"])
def test_wellposed_vs_numpy_solve(name):
    """tests/test_well_posed.py:31-52: 3x3 fp64, cond < 1000, agreement with numpy at 1e-10."""
    rng = np.random.default_rng(hash(name) % 1000)
    spd = name in ("cholesky", "cg")
    for _ in range(5):
        a = rand_cond(rng, 3, spd=spd)
        if name == "tridiagonal":
            a = np.triu(np.tril(a, 1), -1)
            if np.linalg.cond(a) > 1000:
                continue
        b = a @ rng.standard_normal(3)
        ref = np.linalg.solve(a, b)
        if name == "lu":
            x = oracle.lu_compute(oracle.lu_init(a), b)
            xt = oracle.lu_compute(oracle.lu_init(a), b, trans=1)
            assert np.allclose(xt, np.linalg.solve(a.T, b), atol=1e-10)
        elif name == "qr":
            x = oracle.qr_compute(oracle.qr_init(a), b)
        elif name == "cholesky":
            x = oracle.cholesky_compute(oracle.cholesky_init(a), b)
            xn = oracle.cholesky_compute(oracle.cholesky_init(-a, is_nsd=True), b)
            assert np.allclose(xn, -ref, atol=1e-10)
        elif name == "tridiagonal":
            x = oracle.tridiagonal_compute(np.diag(a), np.diag(a, -1), np.diag(a, 1), b)
        else:
            x, res, st = getattr(oracle, name)(a, b, 1e-12, 1e-12)
            assert res == RESULTS.successful, (name, res)
        assert np.allclose(x, ref, atol=1e-9, rtol=1e-9), name


def test_nonsquare_vs_lstsq():
    """tests/test_singular.py:102-241: QR (tall + wide) and LSMR vs lstsq."""
    rng = np.random.default_rng(4)
    for shape in ((5, 3), (3, 5), (2, 3), (3, 2), (40, 7)):
        a = rng.standard_normal(shape)
        b = rng.standard_normal(shape[0])
        ref = np.linalg.lstsq(a, b, rcond=None)[0]
        assert np.allclose(oracle.qr_compute(oracle.qr_init(a), b), ref, atol=1e-10)
        x, res, st = oracle.lsmr(a, b, 1e-12, 1e-12)
        assert res == RESULTS.successful and np.allclose(x, ref, atol=1e-8)


def test_lsmr_matches_scipy_lsmr():
    """lineax's LSMR is a port of SciPy's (lsmr.py:1-34): same iterates and stop codes."""
    from scipy.sparse.linalg import lsmr as sp_lsmr

    rng = np.random.default_rng(5)
    for shape in ((30, 10), (10, 30), (25, 25)):
        a = rng.standard_normal(shape) / 5
        b = rng.standard_normal(shape[0])
        x, res, st = oracle.lsmr(a, b, 1e-10, 1e-10)
        out = sp_lsmr(a, b, atol=1e-10, btol=1e-10, conlim=1e8, maxiter=10 * min(shape))
        assert abs(st["num_steps"] - out[2]) <= 1 and st["istop"] == out[1]
        assert np.allclose(x, out[0], atol=1e-8)


def test_c_lu_oracle_matches_lapack():
    """oracle/getf2.c vs LAPACK getrf/getrs (scipy): identical pivots, solutions within rounding."""
    for dtype, tol in ((np.float32, 2e-3), (np.float64, 1e-10)):
        a, b, _ = gen.gaussian_systems(3, 200, 32, dtype)
        x, lu, piv = clib.lu_factor_solve(a, b)
        for i in range(200):
            lu_ref, piv_ref = oracle.lu_init(a[i])
            assert np.array_equal(piv[i], piv_ref)
            xr = oracle.lu_compute((lu_ref, piv_ref), b[i])
            assert np.max(np.abs(x[i] - xr)) <= tol * max(1.0, np.max(np.abs(xr)))
        xt = clib.lu_solve(lu, piv, b, trans=1)
        for i in range(0, 200, 20):
            assert np.allclose(xt[i], oracle.lu_compute(oracle.lu_init(a[i]), b[i], trans=1),
                               atol=tol * max(1.0, np.abs(xt[i]).max()))


def test_generator_iteration_counts():
    """SURVEY 8(d) probes: easy generator n=256 f32 -> 6 CG steps; n=1024 f64 tol 1e-12 -> 9."""
    a, b, _ = gen.easy_problem(0, 256, np.float32, True)
    assert oracle.cg(a, b, 1e-6, 1e-6)[2]["num_steps"] == 6
    a, b, _ = gen.easy_problem(1, 1024, np.float64, True)
    assert abs(oracle.cg(a, b, 1e-12, 1e-12)[2]["num_steps"] - 9) <= 1
    a, b, _ = gen.easy_problem(2, 512, np.float32, False)
    assert oracle.gmres(a, b, 1e-6, 1e-6)[2]["num_steps"] in (3, 4)
