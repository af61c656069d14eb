"""Oracle-backed CPU kernels for the `lineax_b200::*` custom ops -- TEST DOUBLES ONLY.

The product registers CUDA kernels only and fails loudly on CPU tensors.  To exercise the host
logic (dispatch, PyTree packing, vmap batching rules, result rewriting, sharding) on the GPU-less
CI box, the tests install these CPU implementations, which simply call the oracle."""
import numpy as np
import torch

import oracle
from oracle import clib

_installed = False


def _np(t):
    return t.detach().cpu().numpy()


def _bcast(*pairs):
    """Broadcast batch dims; returns list of arrays reshaped to [B, *core] and the batch shape."""
    shapes = [tuple(t.shape[: t.ndim - c]) for t, c in pairs]
    full = tuple(torch.broadcast_shapes(*shapes))
    outs = []
    for t, c in pairs:
        core = tuple(t.shape[t.ndim - c:])
        outs.append(_np(t.expand(full + core)).reshape((-1,) + core))
    return outs, full


def install():
    global _installed
    if _installed:
        return
    _installed = True
    import lineax_b200._ops  # noqa: F401  (defines the ops)

    reg = lambda name: torch.library.register_kernel(f"lineax_b200::{name}", "cpu")

    @reg("lu_factor")
    def _(a):
        (A,), full = _bcast((a, 2))
        lu, piv = clib.lu_factor(A, threads=1)
        n = a.shape[-1]
        return torch.as_tensor(lu).reshape(full + (n, n)), torch.as_tensor(piv).reshape(full + (n,))

    @reg("lu_solve")
    def _(lu, piv, b, trans):
        (L, P, B), full = _bcast((lu, 2), (piv, 1), (b.to(lu.dtype), 1))
        x = clib.lu_solve(L, P, B, trans=int(trans), threads=1)
        return torch.as_tensor(x).reshape(full + (lu.shape[-1],))

    @reg("lu_factor_solve")
    def _(a, b, keep_state):
        (A, B), full = _bcast((a, 2), (b.to(a.dtype), 1))
        n = a.shape[-1]
        x, lu, piv = clib.lu_factor_solve(A, B, threads=1)
        if not keep_state:
            lu, piv = np.zeros((len(A), 0, 0), A.dtype), np.zeros((len(A), 0), np.int32)
            return (torch.as_tensor(x).reshape(full + (n,)), torch.as_tensor(lu).reshape(full + (0, 0)),
                    torch.as_tensor(piv).reshape(full + (0,)))
        return (torch.as_tensor(x).reshape(full + (n,)), torch.as_tensor(lu).reshape(full + (n, n)),
                torch.as_tensor(piv).reshape(full + (n,)))

    @reg("cg")
    def _(a, b, precond, y0, rtol, atol, max_steps, stabilise_every, flags):
        pairs = [(a, 2), (b.to(a.dtype), 1)]
        if precond is not None:
            pairs.append((precond, 2))
        if y0 is not None:
            pairs.append((y0, 1))
        arrs, full = _bcast(*pairs)
        A, B = arrs[0], arrs[1]
        M = arrs[2] if precond is not None else None
        Y = arrs[-1] if y0 is not None else None
        xs, rs, ss = [], [], []
        for i in range(len(A)):
            x, r, st = oracle.cg(A[i], B[i], rtol, atol, y0=None if Y is None else Y[i],
                                 preconditioner=None if M is None else M[i],
                                 max_steps=max_steps if flags & 4 else None,
                                 stabilise_every=None if stabilise_every == 0 else stabilise_every,
                                 is_nsd=bool(flags & 2))
            xs.append(x), rs.append(r), ss.append(st["num_steps"])
        n = a.shape[-1]
        return (torch.as_tensor(np.stack(xs)).reshape(full + (n,)),
                torch.as_tensor(np.array(rs, np.int32)).reshape(full),
                torch.as_tensor(np.array(ss, np.int32)).reshape(full))

    @reg("postprocess")
    def _(x, b, result):
        (X, B, R), full = _bcast((x, 1), (b, 1), (result.to(torch.int32), 0))
        out = np.array([oracle.postprocess(X[i], int(R[i]), B[i]) for i in range(len(X))], np.int32)
        return torch.as_tensor(out).reshape(full)

    @reg("matvec")
    def _(a, x, trans):
        dt = torch.promote_types(a.dtype, x.dtype)
        (A, X), full = _bcast((a.to(dt), 2), (x.to(dt), 1))
        y = np.einsum("bji,bj->bi" if trans else "bij,bj->bi", A, X)
        return torch.as_tensor(y).reshape(full + (y.shape[-1],))

    @reg("diagonal_solve")
    def _(d, b, rcond):
        dt = torch.promote_types(d.dtype, b.dtype)
        (D, B), full = _bcast((d.to(dt), 1), (b.to(dt), 1))
        x = np.stack([oracle.diagonal_compute(D[i], B[i], well_posed=rcond < 0, rcond=None if rcond < 0 else rcond)
                      for i in range(len(D))])
        return torch.as_tensor(x).reshape(full + (d.shape[-1],))

    @reg("qr_factor")
    def _(a):
        (A,), full = _bcast((a, 2))
        outs = [oracle.qr_init(A[i])[0] for i in range(len(A))]
        aq = np.stack([o[0] for o in outs])
        taus = np.stack([o[1] for o in outs])
        return (torch.as_tensor(aq).reshape(full + aq.shape[1:]), torch.as_tensor(taus).reshape(full + taus.shape[1:]))

    @reg("qr_solve")
    def _(a, taus, b, trans):
        (A, T, B), full = _bcast((a, 2), (taus, 1), (b.to(a.dtype), 1))
        x = np.stack([oracle.qr_compute(((A[i], T[i]), bool(trans)), B[i]) for i in range(len(A))])
        return torch.as_tensor(x).reshape(full + (x.shape[-1],))

    @reg("qr_apply_qt")
    def _(a, taus, b):
        from scipy.linalg import get_lapack_funcs

        (A, T, B), full = _bcast((a, 2), (taus, 1), (b.to(a.dtype), 1))
        outs = []
        for i in range(len(A)):
            (ormqr,) = get_lapack_funcs(("ormqr",), (A[i],))
            c = np.asfortranarray(B[i].reshape(-1, 1))
            q, _, info = ormqr("L", "T", np.asfortranarray(A[i]), T[i], c, max(1, 64 * A[i].shape[0]))
            outs.append(q[: A[i].shape[1], 0])
        y = np.stack(outs)
        return torch.as_tensor(y).reshape(full + (y.shape[-1],))

    @reg("tridiagonal_solve")
    def _(d, dl, du, b):
        (D, L, U, B), full = _bcast((d, 1), (dl, 1), (du, 1), (b.to(d.dtype), 1))
        x = np.stack([oracle.tridiagonal_compute(D[i], L[i], U[i], B[i]) for i in range(len(D))])
        return torch.as_tensor(x).reshape(full + (d.shape[-1],))
