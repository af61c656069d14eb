"""Shared helpers for the parity tests (NumPy <-> CUDA tensors, tolerances)."""
import numpy as np
import torch

# north_star: solutions agree within 1e-5 relative in fp32 and 1e-12 in fp64, num_steps +-2
RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def rel_err(x, ref):
    """max-norm relative error per system (last axis)."""
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    den = np.maximum(np.max(np.abs(ref), axis=-1), 1e-300)
    return np.max(np.abs(x - ref), axis=-1) / den


def assert_close(x, ref, dtype, factor=1.0, what="solution"):
    x, ref = np.asarray(x), np.asarray(ref)
    # non-finite entries must coincide (e.g. both implementations run into the same 0/0)
    assert np.array_equal(np.isfinite(x), np.isfinite(ref)), f"{what}: non-finite patterns differ"
    fin = np.isfinite(ref)
    x, ref = np.where(fin, x, 0), np.where(fin, ref, 0)
    err = np.max(rel_err(x, ref)) if x.size else 0.0
    tol = RTOL[np.dtype(dtype)] * factor
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol:.1e}"


EPS = {np.dtype(np.float32): float(np.finfo(np.float32).eps), np.dtype(np.float64): float(np.finfo(np.float64).eps)}


def cond_inf(a):
    """Infinity-norm condition number of a (square) matrix, computed in float64."""
    return float(np.linalg.cond(np.asarray(a, dtype=np.float64), np.inf))


def cond_2(a):
    """2-norm condition number (works for rectangular matrices)."""
    return float(np.linalg.cond(np.asarray(a, dtype=np.float64)))


def tol_for(dtype, cond=1.0, c=8.0, solver_tol=0.0):
    """Tolerance of a parity check against the oracle.

    Base = north_star (1e-5 relative in fp32, 1e-12 in fp64).  It is widened ONLY by what the instance
    itself makes legitimate: two backward-stable computations of the same solution may differ by
    `c * cond * eps`, and two Krylov runs that stop one step apart (num_steps is allowed +-2) differ by
    up to the stopping threshold times the condition number (`10 * solver_tol * cond`).  For the
    BASELINE generators (cond of a few units) this is the flat north_star tolerance."""
    dt = np.dtype(dtype)
    return max(RTOL[dt], c * cond * EPS[dt], 10.0 * solver_tol * cond)


def assert_close_tol(x, ref, tol, what="solution"):
    x, ref = np.asarray(x), np.asarray(ref)
    assert np.array_equal(np.isfinite(x), np.isfinite(ref)), f"{what}: non-finite patterns differ"
    fin = np.isfinite(ref)
    x, ref = np.where(fin, x, 0), np.where(fin, ref, 0)
    err = float(np.max(rel_err(x, ref))) if x.size else 0.0
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol:.1e}"
    return err


def backward_error(a, x, b):
    """Normwise backward error ||Ax - b||_inf / (||A||_inf ||x||_inf + ||b||_inf), float64 arithmetic."""
    a, x, b = (np.asarray(v, dtype=np.float64) for v in (a, x, b))
    r = a @ x - b
    den = np.abs(a).sum(-1).max() * np.abs(x).max() + np.abs(b).max()
    return float(np.abs(r).max() / max(den, 1e-300))


def max_cond(a):
    """Largest 2-norm condition number over a batch of matrices (float64)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2:
        return cond_2(a)
    return max(cond_2(m) for m in a)
