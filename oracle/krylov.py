"""Krylov drivers restated in NumPy (test infrastructure, see oracle/__init__.py).

Each function follows the reference line by line:
  cg        -> lineax/_solver/cg.py:114-227
  bicgstab  -> lineax/_solver/bicgstab.py:78-205
  gmres     -> lineax/_solver/gmres.py:106-413
  lsmr      -> lineax/_solver/lsmr.py:94-409
All arithmetic is done in the dtype of the inputs (np.float32 / np.float64);
Python-float tolerances are "weak" scalars exactly as in JAX.

Operators are given either as a dense 2-D array or as an object with
`.mv(x)` / `.rmv(x)` (transpose-conjugate product).  Every function returns
`(solution, result_code, stats_dict)` BEFORE the `_solve.py:104-123`
post-processing (apply `oracle.postprocess` for that).
"""
import numpy as np

from .results import RESULTS
from .solve import max_norm, two_norm, tree_dot, resolve_rcond


class _Dense:
    def __init__(self, a):
        self.a = a

    def mv(self, x):
        return self.a @ x

    def rmv(self, x):
        return self.a.T @ x


def _op(a):
    return _Dense(a) if isinstance(a, np.ndarray) else a


def _prec(M):
    if M is None:
        return lambda r: r
    M = _op(M)
    return M.mv


def _has_scale(rtol, atol):
    # cg.py:140-145 -- static Python-scalar test
    return not (rtol == 0 and atol == 0)


def _not_converged_factory(has_scale, rtol, atol, b, norm):
    dt = b.dtype.type
    if has_scale:
        b_scale = dt(atol) + dt(rtol) * np.abs(b)  # cg.py:146-147

    def not_converged(r, diff, y):
        # cg.py:149-160 (identical in bicgstab.py:115-126, gmres.py:130-141)
        if not has_scale:
            return True
        with np.errstate(all="ignore"):
            y_scale = dt(atol) + dt(rtol) * np.abs(y)
            norm1 = norm(r / b_scale)
            norm2 = norm(diff / y_scale)
        return bool(norm1 > 1) | bool(norm2 > 1)

    return not_converged


def _final_result(num_steps, max_steps_arg, max_steps, has_scale):
    # cg.py:213-222
    if max_steps_arg is None:
        return RESULTS.singular if num_steps == max_steps else RESULTS.successful
    elif has_scale:
        return RESULTS.max_steps_reached if num_steps == max_steps else RESULTS.successful
    return RESULTS.successful


def cg(A, b, rtol, atol, *, y0=None, preconditioner=None, max_steps=None,
       stabilise_every=10, is_nsd=False, norm=max_norm):
    """Preconditioned CG exactly as lineax/_solver/cg.py:114-227."""
    op = _op(A)
    b = np.asarray(b)
    dt = b.dtype.type
    sign = dt(-1) if is_nsd else dt(1)  # cg.py:100-101: operator = -operator
    mv = lambda x: sign * op.mv(x)
    M = _prec(preconditioner)
    size = b.size
    ms = 10 * size if max_steps is None else max_steps  # cg.py:124-127
    y = np.zeros_like(b) if y0 is None else np.array(y0, dtype=b.dtype)
    with np.errstate(all="ignore"):
        r = b - mv(y)  # cg.py:128 (always evaluated)
        p = M(r)
        gamma = tree_dot(p, r)
        rcond = resolve_rcond(None, size, size, b.dtype)  # cg.py:131
        diff = np.full_like(y, np.inf)
        step = 0
        has_scale = _has_scale(rtol, atol)
        not_converged = _not_converged_factory(has_scale, rtol, atol, b, norm)
        while bool(gamma > 0) and step < ms and not_converged(r, diff, y):  # cg.py:162-167
            mat_p = mv(p)
            inner_prod = tree_dot(mat_p, p)
            alpha = gamma / inner_prod
            if not (np.abs(inner_prod) > dt(100) * rcond * np.abs(gamma)):  # cg.py:174-178
                alpha = dt(np.nan)
            diff = alpha * p
            y = y + diff
            step += 1
            if stabilise_every == 1:
                r = b - mv(y)
            elif stabilise_every is None:
                r = r - alpha * mat_p
            elif step % stabilise_every == 0:  # cg.py:187-200
                r = b - mv(y)
            else:
                r = r - alpha * mat_p
            z = M(r)
            gamma_prev = gamma
            gamma = tree_dot(z, r)
            beta = gamma / gamma_prev
            p = z + beta * p
    result = _final_result(step, max_steps, ms, has_scale)
    if is_nsd:
        y = -y  # cg.py:224-225
    return y, result, {"num_steps": step, "max_steps": max_steps}


def bicgstab(A, b, rtol, atol, *, y0=None, preconditioner=None, max_steps=None,
             x64=None, norm=max_norm):
    """Right-preconditioned BiCGStab exactly as lineax/_solver/bicgstab.py:78-205.

    `x64` mirrors `jax.config.jax_enable_x64` (bicgstab.py:110): True -> `== 0`
    breakdown test, False -> signed `< 1e-16` test.  None -> dtype is float64.
    """
    op = _op(A)
    b = np.asarray(b)
    dt = b.dtype.type
    if x64 is None:
        x64 = b.dtype == np.float64
    M = _prec(preconditioner)
    size = b.size
    ms = 10 * size if max_steps is None else max_steps
    has_scale = _has_scale(rtol, atol)
    not_converged = _not_converged_factory(has_scale, rtol, atol, b, norm)
    y = np.zeros_like(b) if y0 is None else np.array(y0, dtype=b.dtype)

    def breakdown_occurred(omega, alpha, rho):  # bicgstab.py:107-113
        if x64:
            return bool(omega == 0.0) | bool(alpha == 0.0) | bool(rho == 0.0)
        t = dt(1e-16)
        return bool(omega < t) | bool(alpha < t) | bool(rho < t)

    with np.errstate(all="ignore"):
        r0 = b - op.mv(y)
        r = r0
        alpha = omega = rho = dt(1.0)
        p = np.zeros_like(b)
        v = np.zeros_like(b)
        diff = np.full_like(y, np.inf)
        step = 0
        while (not breakdown_occurred(omega, alpha, rho)) and not_converged(r, diff, y) and step < ms:
            rho_new = tree_dot(r0, r)
            beta = (rho_new / rho) * (alpha / omega)
            p = r + beta * (p - omega * v)
            x = M(p)
            v = op.mv(x)
            alpha = rho_new / tree_dot(r0, v)
            s = r - alpha * v
            z = M(s)
            t = op.mv(z)
            omega = tree_dot(s, t) / tree_dot(t, t)
            diff = alpha * x + omega * z
            y = y + diff
            r = s - omega * t
            rho = rho_new
            step += 1
        result = _final_result(step, max_steps, ms, has_scale)
        if breakdown_occurred(omega, alpha, rho) and not_converged(r, diff, y):  # bicgstab.py:199-202
            result = RESULTS.breakdown
    return y, result, {"num_steps": step, "max_steps": max_steps}


def _normalise(x, eps):
    # gmres.py:401-413
    nrm = two_norm(x)
    if eps is None:
        eps = np.finfo(nrm.dtype).eps
    eps = nrm.dtype.type(eps)
    breakdown = bool(nrm < eps)
    safe = nrm.dtype.type(np.inf) if breakdown else nrm
    return x / safe, nrm, breakdown


def gmres(A, b, rtol, atol, *, y0=None, preconditioner=None, max_steps=None,
          restart=20, stagnation_iters=20, norm=max_norm):
    """Restarted GMRES exactly as lineax/_solver/gmres.py:106-413."""
    from .direct import qr_init, qr_compute

    op = _op(A)
    b = np.asarray(b)
    dt = b.dtype.type
    M = _prec(preconditioner)
    size = b.size
    ms = 10 * size if max_steps is None else max_steps
    restart = min(restart, size)  # gmres.py:128
    has_scale = _has_scale(rtol, atol)
    not_converged = _not_converged_factory(has_scale, rtol, atol, b, norm)
    eps = np.finfo(b.dtype).eps

    def main_gmres(y, r):  # gmres.py:254-310
        r_normalised, r_norm, initial_breakdown = _normalise(r, None)
        basis = np.zeros((size, restart + 1), dtype=b.dtype)
        basis[:, 0] = r_normalised
        coeff = np.eye(restart, restart + 1, dtype=b.dtype)
        breakdown = initial_breakdown
        k = 0
        while k < restart and not breakdown:  # gmres.py:269-271
            # _arnoldi_gram_schmidt, gmres.py:331-399
            w = M(op.mv(basis[:, k]))
            step_norm = two_norm(w)
            proj = basis.T @ w  # all restart+1 columns (zeros beyond k)
            w = w - basis @ proj
            w_n, step_norm_new, breakdown = _normalise(w, step_norm * dt(eps))
            basis[:, k + 1] = w_n
            proj[k + 1] = step_norm_new
            coeff[k, :] = proj  # pred = ~carry breakdown, which is False inside the loop
            k += 1
        beta_vec = np.zeros(restart + 1, dtype=b.dtype)
        beta_vec[0] = r_norm
        # gmres.py:301-303: linear_solve(MatrixLinearOperator(coeff.T), beta_vec, QR(), throw=False)
        z = qr_compute(qr_init(np.ascontiguousarray(coeff.T)), beta_vec)
        diff = basis[:, :-1] @ z
        return y + diff, diff, breakdown

    with np.errstate(all="ignore"):
        y = np.zeros_like(b) if y0 is None else np.array(y0, dtype=b.dtype)
        r = np.zeros_like(b)  # dummy residual, gmres.py:195
        breakdown = False
        deferred = False
        diff = np.full_like(y, np.inf)
        r_min = dt(np.inf)
        step = 0
        stag = 0
        while ((not deferred) and stag < stagnation_iters and not_converged(r, diff, y)
               and step < ms) or step == 0:  # gmres.py:143-158
            if step == 0:  # first_gmres, gmres.py:313-314
                y_new, diff_new, bd = y, np.full_like(y, np.inf), False
            else:
                y_new, diff_new, bd = main_gmres(y, r)
            r_new = M(b - op.mv(y_new))  # gmres.py:318
            r_new_norm = norm(r_new)
            r_decreased = bool((r_new_norm - r_min) < 0)  # gmres.py:175-179
            stag = 0 if r_decreased else stag + 1
            r_min = np.minimum(r_new_norm, r_min)
            y, r, deferred, breakdown, diff = y_new, r_new, breakdown, bd, diff_new
            step += 1
        result = _final_result(step, max_steps, ms, has_scale)
        if stag >= stagnation_iters:  # gmres.py:228-230
            result = RESULTS.stagnation
        if deferred and not_converged(r, diff, y):  # gmres.py:235-237
            result = RESULTS.breakdown
    return y, result, {"num_steps": step, "max_steps": max_steps}


def _givens(a, b):
    """Stable Givens rotation, lineax/_solver/lsmr.py:361-409 (real a, b)."""
    dt = type(a)
    if a == 0 or b == 0:
        if b == 0:
            return np.sign(a), dt(0.0), np.abs(a)
        return dt(0.0), np.sign(b), np.abs(b)
    if np.abs(b) > np.abs(a):
        tau = a / b
        s = np.sign(b) / np.sqrt(dt(1.0) + tau * tau)
        c = s * tau
        r = b / (dt(1.0) if s == 0 else s)
        return c, s, r
    tau = b / a
    c = np.sign(a) / np.sqrt(dt(1.0) + tau * tau)
    s = c * tau
    r = a / (dt(1.0) if c == 0 else c)
    return c, s, r


def lsmr(A, b, rtol, atol, *, y0=None, max_steps=None, conlim=1e8, norm=two_norm):
    """LSMR (damp = 0) exactly as lineax/_solver/lsmr.py:94-359."""
    op = _op(A)
    b = np.asarray(b)
    dtype = b.dtype
    dt = dtype.type
    if isinstance(A, np.ndarray):
        m, n = A.shape
    else:
        m, n = op.shape
    has_scale = _has_scale(rtol, atol)
    min_dim = min(m, n)
    if max_steps is None:  # lsmr.py:121-129
        imax = np.iinfo(np.dtype(f"int{dtype.itemsize * 8}")).max
        ms = imax if min_dim > imax / 10 else min_dim * 10
    else:
        ms = max_steps
    with np.errstate(all="ignore"):
        x = np.zeros(n, dtype=dtype) if y0 is None else np.array(y0, dtype=dtype)
        u = b - op.mv(x)
        normb = norm(b)
        beta = norm(u)
        if beta == 0:  # lsmr.py:143-151
            v = np.zeros(n, dtype=dtype)
            alpha = dt(0.0)
        else:
            u = u / beta
            v = op.rmv(u)
            alpha = norm(v)
        v = v / (dt(1.0) if alpha == 0 else alpha)
        h = v.copy()
        hbar = np.zeros(n, dtype=dtype)
        itn = 0
        zetabar = alpha * beta
        alphabar = alpha
        rho = dt(1.0); rhobar = dt(1.0); cbar = dt(1.0); sbar = dt(0.0)
        betadd = beta; betad = dt(0.0); rhodold = dt(1.0); tautildeold = dt(0.0)
        thetatilde = dt(0.0); zeta = dt(0.0); delta = dt(0.0)
        normA2 = alpha * alpha; maxrbar = dt(0.0); minrbar = dt(np.finfo(dtype).max)
        condA = dt(1.0)
        istop = 0
        normr = beta
        normAr = alpha * beta
        if alpha == 0:
            istop = 2
        if beta == 0:
            istop = 1
        damp = dt(0.0)
        while istop == 0:
            itn += 1
            u = u * -alpha
            u = u + op.mv(v)
            beta = norm(u)
            if beta != 0:  # lsmr.py:218-236
                u = u / beta
                v = v * -beta
                v = v + op.rmv(u)
                alpha = norm(v)
                v = v / (dt(1.0) if alpha == 0 else alpha)
            chat, shat, alphahat = _givens(alphabar, damp)
            rhoold = rho
            c, s, rho = _givens(alphahat, beta)
            thetanew = s * alpha
            alphabar = c * alpha
            rhobarold = rhobar
            zetaold = zeta
            thetabar = sbar * rho
            rhotemp = cbar * rho
            cbar, sbar, rhobar = _givens(cbar * rho, thetanew)
            zeta = cbar * zetabar
            zetabar = -sbar * zetabar
            hbar = hbar * -(thetabar * rho / (rhoold * rhobarold))
            hbar = hbar + h
            x = x + (zeta / (rho * rhobar)) * hbar
            h = h * -(thetanew / rho)
            h = h + v
            betaacute = chat * betadd
            betacheck = -shat * betadd
            betahat = c * betaacute
            betadd = -s * betaacute
            thetatildeold = thetatilde
            ctildeold, stildeold, rhotildeold = _givens(rhodold, thetabar)
            thetatilde = stildeold * rhobar  # lsmr.py:286 reads the UPDATED rhobar
            rhodold = ctildeold * rhobar
            betad = -stildeold * betad + ctildeold * betahat
            tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold
            taud = (zeta - thetatilde * tautildeold) / rhodold
            delta = delta + betacheck * betacheck
            normr = np.sqrt(delta + (betad - taud) ** 2 + betadd * betadd)
            normA2 = normA2 + beta * beta
            normA = np.sqrt(normA2)
            normA2 = normA2 + alpha * alpha
            maxrbar = np.maximum(maxrbar, rhobarold)
            if itn > 1:
                minrbar = np.minimum(minrbar, rhobarold)
            condA = np.maximum(maxrbar, rhotemp) / np.minimum(minrbar, rhotemp)
            normAr = np.abs(zetabar)
            normx = norm(x)
            well_posed_tol = dt(atol) + dt(rtol) * (normA * normx + normb)
            least_squares_tol = dt(atol) + dt(rtol) * (normA * normr)
            if itn >= ms:  # lsmr.py:317-329, in overwrite order
                istop = 4
            if condA > conlim:
                istop = 3
            if normAr < least_squares_tol:
                istop = 2
            if normr < well_posed_tol:
                istop = 1
        stats = {
            "num_steps": itn, "istop": istop, "norm_r": normr, "norm_Ar": normAr,
            "norm_A": np.sqrt(normA2), "cond_A": condA, "norm_x": norm(x),
        }
    result = _final_result(itn, max_steps, ms, has_scale)
    if istop < 3:  # lsmr.py:356-357
        result = RESULTS.successful
    if istop == 3:
        result = RESULTS.conlim
    return x, result, stats
