"""Direct solvers restated on SciPy LAPACK (test infrastructure, see oracle/__init__.py).

  lu_*          -> lineax/_solver/lu.py:43-66          (getrf / getrs)
  cholesky_*    -> lineax/_solver/cholesky.py:43-78    (potrf upper / potrs)
  qr_*          -> lineax/_solver/qr.py:55-94          (geqrf / ormqr / trtrs)
  tridiagonal_* -> lineax/_solver/tridiagonal.py:54-72 (gtsv, partial pivoting)
  diagonal_*    -> lineax/_solver/diagonal.py:66-83
  triangular_*  -> lineax/_solver/triangular.py:68-85

jaxlib's CPU backend lowers these to the same LAPACK routines (unverifiable
here: jax is not installed); SciPy's OpenBLAS build is the stand-in.
All functions take/return NumPy arrays and return only the solution: these
solvers always report RESULTS.successful before `_solve.py:104-123`.
"""
import numpy as np
from scipy.linalg import get_lapack_funcs

from .solve import resolve_rcond


def _lapack(names, a):
    return get_lapack_funcs(names, (a,))


def lu_init(A):
    """lu.py:43-54 -> (lu, piv[int32, 0-based row-swap sequence]); LAPACK getrf."""
    A = np.asarray(A)
    (getrf,) = _lapack(("getrf",), A)
    lu, piv, info = getrf(A)
    assert info >= 0
    return np.ascontiguousarray(lu), piv.astype(np.int32)


def lu_compute(state, b, trans=0):
    """lu.py:56-66: lu_solve((lu, piv), b, trans)."""
    lu, piv = state
    (getrs,) = _lapack(("getrs",), lu)
    x, info = getrs(lu, piv, np.asarray(b, dtype=lu.dtype), trans=trans)
    assert info == 0
    return x


def cholesky_init(A, is_nsd=False):
    """cholesky.py:43-63: upper factor U with (±A) = U^T U; non-PD input gives NaN like XLA's potrf."""
    A = np.asarray(A)
    if is_nsd:
        A = -A
    (potrf,) = _lapack(("potrf",), A)
    c, info = potrf(A, lower=0, clean=1)
    if info != 0:
        c = np.full_like(A, np.nan)
    return np.ascontiguousarray(c), is_nsd


def cholesky_compute(state, b):
    """cholesky.py:65-78."""
    c, is_nsd = state
    (potrs,) = _lapack(("potrs",), c)
    x, info = potrs(c, np.asarray(b, dtype=c.dtype), lower=0)
    if is_nsd:
        x = -x
    return x


def qr_init(A):
    """qr.py:55-65: geqrf of A (or of A^T when wide) -> ((a, taus), transpose)."""
    A = np.asarray(A)
    m, n = A.shape
    transpose = n > m
    if transpose:
        A = A.T
    geqrf, geqrf_lwork = _lapack(("geqrf", "geqrf_lwork"), A)
    lwork, _ = geqrf_lwork(*A.shape)  # optimal workspace: LAPACK's blocked algorithm, as jaxlib queries it
    a, taus, _, info = geqrf(np.asfortranarray(A), lwork=int(lwork))
    assert info == 0
    return (np.ascontiguousarray(a), taus), transpose


def qr_compute(state, b):
    """qr.py:67-94: tall -> R^{-1} (Q^H b)[:n]; wide -> Q [R^{-T} b; 0] (minimum norm)."""
    (a, taus), transpose = state
    n_full, n_min = a.shape
    ormqr, trtrs = _lapack(("ormqr", "trtrs"), a)
    b = np.asarray(b, dtype=a.dtype)
    r = a[:n_min]
    af = np.asfortranarray(a)
    lwork = max(1, 64 * max(n_full, n_min))
    if transpose:
        y, info = trtrs(r, b, lower=0, trans=1)
        y_pad = np.zeros((n_full, 1), dtype=a.dtype, order="F")
        y_pad[:n_min, 0] = y
        # Q.conj() @ z  ==  Q @ z for real dtypes
        out, _, info = ormqr("L", "N", af, taus, y_pad, lwork)
        return np.ascontiguousarray(out[:, 0])
    c = np.asfortranarray(b.reshape(-1, 1))
    qhb, _, info = ormqr("L", "T", af, taus, c, lwork)
    x, info = trtrs(r, qhb[:n_min, 0], lower=0, trans=0)
    return x


def tridiagonal_compute(diagonal, lower, upper, b):
    """tridiagonal.py:54-72: LAPACK gtsv (Gaussian elimination with partial pivoting)."""
    diagonal = np.asarray(diagonal)
    (gtsv,) = _lapack(("gtsv",), diagonal)
    if diagonal.size == 1:
        return np.asarray(b, dtype=diagonal.dtype) / diagonal
    _, _, _, x, info = gtsv(lower, diagonal, upper, np.asarray(b, dtype=diagonal.dtype))
    if info > 0:  # exactly singular U: LAPACK stops; XLA's kernels produce inf/nan instead
        x = np.full_like(x, np.nan)
    return x.reshape(-1)


def diagonal_compute(diag, b, well_posed=False, rcond=None):
    """diagonal.py:66-83: well-posed -> b / d; else pseudo-inverse with rcond masking."""
    diag = np.asarray(diag)
    b = np.asarray(b, dtype=diag.dtype)
    with np.errstate(all="ignore"):
        if well_posed:
            return b / diag
        size = diag.size
        rc = resolve_rcond(rcond, size, size, diag.dtype)
        abs_diag = np.abs(diag)
        mask = abs_diag > rc * (np.max(abs_diag) if size else diag.dtype.type(0))
        return b / np.where(mask, diag, diag.dtype.type(np.inf))  # diagonal.py:78-80


def triangular_compute(A, b, lower, unit_diagonal=False, trans=0):
    """triangular.py:68-85: solve_triangular."""
    A = np.asarray(A)
    (trtrs,) = _lapack(("trtrs",), A)
    x, info = trtrs(A, np.asarray(b, dtype=A.dtype), lower=int(lower), trans=trans,
                    unitdiag=int(unit_diagonal))
    return x
