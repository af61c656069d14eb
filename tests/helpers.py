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
