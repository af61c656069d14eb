"""Seeded synthetic inputs shared by tests and bench (test infrastructure, see oracle/__init__.py).

Generators follow SURVEY.md section 8(d); the "easy" generator is the reference's own
`create_easy_iterative_problem` (benchmarks/solver_speeds.py:146-152):
M = N(0,1)/n + 2 I, SPD case A = M^T M, x* ~ N(0,1), b = A x*.
"""
import numpy as np


def easy_matrix(rng, n, dtype, spd):
    m = rng.standard_normal((n, n)) / n + 2.0 * np.eye(n)
    a = m.T @ m if spd else m
    return a.astype(dtype)


def easy_problem(seed, n, dtype=np.float32, spd=True, batch=None):
    """(A, b, x_true) with optional leading batch dimension."""
    rng = np.random.default_rng(seed)
    if batch is None:
        a = easy_matrix(rng, n, dtype, spd)
        x = rng.standard_normal(n).astype(dtype)
        return a, (a.astype(np.float64) @ x).astype(dtype), x
    a = np.stack([easy_matrix(rng, n, dtype, spd) for _ in range(batch)])
    x = rng.standard_normal((batch, n)).astype(dtype)
    b = np.einsum("bij,bj->bi", a.astype(np.float64), x).astype(dtype)
    return a, b, x


def spectrum_spd(seed, n, cond, dtype=np.float32, batch=None):
    """A = Q diag(geomspace(1, cond, n)) Q^T (SURVEY 8(d), C3 secondary)."""
    rng = np.random.default_rng(seed)

    def one():
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        a = (q * np.geomspace(1.0, cond, n)) @ q.T
        a = 0.5 * (a + a.T)
        return a.astype(dtype)

    if batch is None:
        a = one()
        x = rng.standard_normal(n).astype(dtype)
        return a, (a.astype(np.float64) @ x).astype(dtype), x
    a = np.stack([one() for _ in range(batch)])
    x = rng.standard_normal((batch, n)).astype(dtype)
    b = np.einsum("bij,bj->bi", a.astype(np.float64), x).astype(dtype)
    return a, b, x


def gaussian_systems(seed, batch, n, dtype=np.float32):
    """C2: A ~ N(0,1)^{n x n}, b = A x*."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((batch, n, n)).astype(dtype)
    x = rng.standard_normal((batch, n)).astype(dtype)
    b = np.einsum("bij,bj->bi", a.astype(np.float64), x).astype(dtype)
    return a, b, x


def tall_lstsq(seed, m, n, dtype=np.float32, noise=0.1):
    """C5: A = N(0,1)/sqrt(m), b = A x* + noise N(0,1)."""
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal((m, n)) / np.sqrt(m)).astype(dtype)
    x = rng.standard_normal(n).astype(dtype)
    b = (a.astype(np.float64) @ x + noise * rng.standard_normal(m)).astype(dtype)
    return a, b, x


def tridiagonal_systems(seed, batch, n, dtype=np.float32):
    """C5: strictly diagonally dominant: d = 4 + |N|, |l| + |u| < |d|."""
    rng = np.random.default_rng(seed)
    d = (4.0 + np.abs(rng.standard_normal((batch, n)))).astype(dtype)
    l = rng.standard_normal((batch, n - 1)).astype(dtype)
    u = rng.standard_normal((batch, n - 1)).astype(dtype)
    b = rng.standard_normal((batch, n)).astype(dtype)
    return d, np.clip(l, -1.9, 1.9), np.clip(u, -1.9, 1.9), b


def poisson_matrix(n, dtype=np.float64):
    """tests/helpers.py:101-107."""
    return (-2 * np.eye(n) + np.eye(n, k=1) + np.eye(n, k=-1)).astype(dtype)
