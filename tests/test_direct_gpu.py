"""Cholesky / QR / Tridiagonal / Diagonal / Triangular parity (GPU) vs the SciPy-LAPACK oracle."""
import numpy as np
import pytest

import oracle
from oracle import gen
from tests.helpers import (assert_close, assert_close_tol, backward_error, cond_2, cond_inf, dev, host,
                           tol_for, EPS)

pytestmark = pytest.mark.gpu


def _ops():
    from lineax_b200 import _ops

    return _ops


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 5, 32, 64, 150, 250])
@pytest.mark.parametrize("nsd", [False, True])
def test_cholesky(n, dtype, nsd):
    a, b, _ = gen.easy_problem(n, n, dtype, spd=True, batch=4)
    if nsd:
        a = -a
    f = _ops().cholesky_factor(dev(a), nsd)
    x = _ops().cholesky_solve(f, dev(b), nsd)
    for i in range(4):
        st = oracle.cholesky_init(a[i], is_nsd=nsd)
        # easy SPD generator: cond of a few units -> flat north_star tolerance (1e-5 / 1e-12)
        tol = tol_for(dtype, cond_inf(a[i]))
        assert tol <= 10 * tol_for(dtype), "generator is expected to be well conditioned"
        assert_close_tol(np.triu(host(f)[i]), np.triu(st[0]), tol, "factor")
        assert np.all(np.tril(host(f)[i], -1) == 0)
        assert_close_tol(host(x)[i], oracle.cholesky_compute(st, b[i]), tol)
        assert backward_error(a[i], host(x)[i], b[i]) <= 4 * n * EPS[np.dtype(dtype)]


def test_cholesky_not_pd_gives_nan():
    a = np.array([[[1.0, 2.0], [2.0, 1.0]]], np.float32)
    f = host(_ops().cholesky_factor(dev(a), False))
    assert np.all(np.isnan(f))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(1, 1), (3, 3), (5, 3), (3, 5), (21, 20), (64, 64), (500, 200), (200, 500), (2048, 16)])
def test_qr(shape, dtype):
    m, n = shape
    rng = np.random.default_rng(m * 7 + n)
    a = rng.standard_normal((3, m, n)).astype(dtype)
    b = rng.standard_normal((3, m)).astype(dtype)
    aq, taus = _ops().qr_factor(dev(a))
    x = host(_ops().qr_solve(aq, taus, dev(b), n > m))
    for i in range(3):
        st = oracle.qr_init(a[i])
        xr = oracle.qr_compute(st, b[i])
        # Gaussian inputs are not well conditioned: the legitimate difference between two backward-stable
        # Householder solves is bounded by the instance's own conditioning (kappa for a square system,
        # kappa^2 for a least-squares / minimum-norm problem), nothing else is allowed on top of north_star
        k2 = cond_2(a[i])
        tol = tol_for(dtype, k2 if m == n else k2 * k2)
        assert_close_tol(x[i], xr, tol)
        xl = np.linalg.lstsq(a[i].astype(np.float64), b[i].astype(np.float64), rcond=None)[0]
        assert_close_tol(x[i], xl, tol, "solution vs float64 lstsq")
        (a_ref, taus_ref), _ = st
        rows, cols = a_ref.shape
        # R and taus follow LAPACK's sign conventions (geqr2): compare directly, relative to max |R|
        r_gpu, r_ref = np.triu(host(aq)[i][:cols]), np.triu(a_ref[:cols])
        assert np.max(np.abs(r_gpu - r_ref)) <= tol_for(dtype, k2) * np.max(np.abs(r_ref))
        assert np.max(np.abs(host(taus)[i] - taus_ref)) <= tol_for(dtype, k2)


def test_qr_transposed_state():
    """QR.transpose(): solving A^T x = b reuses the factors of A (qr.py:96-104)."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((1, 6, 4))
    b = rng.standard_normal((1, 4))
    aq, taus = _ops().qr_factor(dev(a))
    x = host(_ops().qr_solve(aq, taus, dev(b), True))[0]  # min-norm solution of A^T x = b
    xl = np.linalg.lstsq(a[0].T, b[0], rcond=None)[0]
    assert np.max(np.abs(x - xl)) < 1e-10


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 512, 700])
def test_tridiagonal_dominant(n, dtype):
    d, l, u, b = gen.tridiagonal_systems(n, 70, n, dtype)
    x = host(_ops().tridiagonal_solve(dev(d), dev(l), dev(u), dev(b)))
    for i in range(0, 70, 9):
        xr = oracle.tridiagonal_compute(d[i], l[i], u[i], b[i])
        assert_close(x[i], xr, dtype)  # diagonally dominant BASELINE generator: flat 1e-5 / 1e-12


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 3, 10, 64, 200])
def test_tridiagonal_general_needs_pivoting(n, dtype):
    """Random (non-dominant) tridiagonals as in tests/helpers.py: pivoting path of gtsv."""
    rng = np.random.default_rng(n)
    d = rng.standard_normal((40, n)).astype(dtype)
    l = rng.standard_normal((40, n - 1)).astype(dtype)
    u = rng.standard_normal((40, n - 1)).astype(dtype)
    b = rng.standard_normal((40, n)).astype(dtype)
    x = host(_ops().tridiagonal_solve(dev(d), dev(l), dev(u), dev(b)))
    for i in range(40):
        T = np.diag(d[i]) + np.diag(l[i], -1) + np.diag(u[i], 1)
        kappa = cond_inf(T)
        if kappa > 1000:
            continue
        xr = oracle.tridiagonal_compute(d[i], l[i], u[i], b[i])
        assert_close_tol(x[i], xr, tol_for(dtype, kappa))
        assert backward_error(T, x[i], b[i]) <= 8 * EPS[np.dtype(dtype)] * 10  # gtsv growth factor slack


def test_tridiagonal_c5_properties():
    """C5 tridiagonal shape at 2^16 systems x 512: residual property."""
    d, l, u, b = gen.tridiagonal_systems(5, 1 << 16, 512, np.float32)
    x = host(_ops().tridiagonal_solve(dev(d), dev(l), dev(u), dev(b)))
    r = d * x - b
    r[:, :-1] += u * x[:, 1:]
    r[:, 1:] += l * x[:, :-1]
    assert np.max(np.abs(r)) < 5e-5
    xr = oracle.tridiagonal_compute(d[12345], l[12345], u[12345], b[12345])
    assert_close(x[12345], xr, np.float32)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_diagonal_and_triangular(dtype):
    rng = np.random.default_rng(0)
    d = rng.standard_normal((6, 17)).astype(dtype)
    d[0, 3] = 0.0
    d[1, :] *= 1e-12
    d[1, 0] = 1.0
    b = rng.standard_normal((6, 17)).astype(dtype)
    eps = np.finfo(dtype).eps
    x = host(_ops().diagonal_solve(dev(d), dev(b), float(2 * eps * 17)))
    xw = host(_ops().diagonal_solve(dev(d), dev(b), -1.0))
    for i in range(6):
        assert_close(x[i], oracle.diagonal_compute(d[i], b[i]), dtype)
        ref = oracle.diagonal_compute(d[i], b[i], well_posed=True)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(xw[i]), fin)
        assert_close(xw[i][fin], ref[fin], dtype)
    for n in (1, 4, 33, 120):
        a = rng.standard_normal((3, n, n)).astype(dtype) + 4 * np.eye(n, dtype=dtype)
        bb = rng.standard_normal((3, n)).astype(dtype)
        for lower in (False, True):
            for unit in (False, True):
                for trans in (False, True):
                    x = host(_ops().triangular_solve(dev(a), dev(bb), lower, unit, trans))
                    for i in range(3):
                        xr = oracle.triangular_compute(a[i], bb[i], lower, unit, int(trans))
                        t = np.tril(a[i]) if lower else np.triu(a[i])
                        if unit:
                            t = t - np.diag(np.diag(t)) + np.eye(n)
                        assert_close_tol(x[i], xr, tol_for(dtype, cond_inf(t)))


@pytest.mark.parametrize("shape,dtype", [((4096, 256), np.float32), ((8192, 130), np.float64),
                                          ((20000, 96), np.float32), ((65536, 64), np.float32),
                                          ((8192, 131), np.float32), ((6000, 67), np.float64),
                                          ((16384, 640), np.float32), ((300000, 64), np.float32),
                                          ((10000, 300), np.float32), ((9000, 132), np.float32)])
def test_qr_large_blocked(shape, dtype):
    """Blocked (compact WY) Householder QR on all SMs for one large tall matrix: same R / taus as
    LAPACK geqrf (identical sign conventions) and the least-squares solution of qr.py:89-92."""
    m, n = shape
    a, b, _ = gen.tall_lstsq(m + n, m, n, dtype)
    aq, taus = _ops().qr_factor(dev(a[None]))
    x = host(_ops().qr_solve(aq, taus, dev(b[None]), False))[0]
    # reference = LAPACK geqrf in FLOAT64 on the same input (identical sign conventions): with m up to
    # 3e5 the float32 LAPACK factors themselves carry ~sqrt(m) eps of rounding, so they cannot arbitrate
    # a 1e-5 comparison; the float64 factors can.  The C5 generator is well conditioned (kappa ~ 1-3).
    (a_ref, taus_ref), _ = oracle.qr_init(a.astype(np.float64))
    k2 = cond_2(a) if m * n <= 1 << 23 else (1 + np.sqrt(n / m)) / (1 - np.sqrt(n / m))
    tol = tol_for(dtype, k2 * k2)
    assert tol <= 3 * tol_for(dtype), "C5 generator is expected to be well conditioned"
    r_gpu, r_ref = np.triu(host(aq)[0][:n]), np.triu(a_ref[:n])
    assert np.max(np.abs(r_gpu - r_ref)) <= tol * np.max(np.abs(r_ref)), "R"
    assert np.max(np.abs(host(taus)[0] - taus_ref)) <= tol, "taus"
    v_gpu, v_ref = np.tril(host(aq)[0], -1), np.tril(a_ref, -1)
    assert np.max(np.abs(v_gpu - v_ref)) <= tol * max(1.0, np.max(np.abs(v_ref))), "Householder vectors"
    xl = np.linalg.lstsq(a.astype(np.float64), b.astype(np.float64), rcond=None)[0]
    assert_close_tol(x, xl, tol, "solution vs float64 lstsq")
    xr = oracle.qr_compute(oracle.qr_init(a), b)
    assert_close_tol(x, xr, tol, "solution vs float32 LAPACK")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [64, 100, 256, 512, 1024])
def test_tridiagonal_mixed_batch_warp_and_pivoting_kernels(n, dtype):
    """A batch that mixes diagonally dominant systems (warp-per-system kernel, csrc/tridiagonal.cu) with
    systems that need gtsv's row interchanges (handed over to the pivoting kernel through the in-kernel
    list), odd batch sizes, and strided operands: every system must match LAPACK gtsv."""
    if dtype == np.float64 and n > 512:
        pytest.skip("fp64 warp kernel covers n <= 512; larger n runs the pivoting kernel (covered elsewhere)")
    rng = np.random.default_rng(n)
    batch = 77
    d, l, u, b = gen.tridiagonal_systems(n, batch, n, dtype)
    hard = rng.random(batch) < 0.4  # these lose dominance: tiny pivots force interchanges
    for i in np.nonzero(hard)[0]:
        k = rng.integers(0, n, size=5)
        d[i, k] = (1e-3 * rng.standard_normal(5)).astype(dtype)
    x = host(_ops().tridiagonal_solve(dev(d), dev(l), dev(u), dev(b)))
    checked = 0
    for i in range(batch):
        T = np.diag(d[i]) + np.diag(l[i], -1) + np.diag(u[i], 1)
        kappa = cond_inf(T)
        if kappa > 1e4:
            continue
        xr = oracle.tridiagonal_compute(d[i], l[i], u[i], b[i])
        assert_close_tol(x[i], xr, tol_for(dtype, kappa), f"system {i} (hard={bool(hard[i])})")
        checked += 1
    assert checked > batch // 2 and hard.any() and (~hard).any()
    # strided views (rows 2 apart): same answers
    xs = host(_ops().tridiagonal_solve(dev(d)[::2], dev(l)[::2], dev(u)[::2], dev(b)[::2]))
    assert np.array_equal(xs, x[::2])
