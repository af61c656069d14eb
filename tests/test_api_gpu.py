"""Public-API tests (GPU), written to read like the reference's own tests:
tests/test_solve.py, test_well_posed.py, test_vmap.py, test_vmap_vmap.py, test_singular.py,
test_lsmr.py, test_transpose.py of patrick-kidger/lineax -- with torch.func.vmap for jax.vmap."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

tol = 1e-12


def _lx():
    import lineax_b200 as lx

    return lx


def t64(x):
    return torch.as_tensor(np.asarray(x, dtype=np.float64)).cuda()


def allclose(a, b, rtol=1e-5, atol=1e-8):
    import torch.utils._pytree as pt

    la, ta = pt.tree_flatten(a)
    lb, tb = pt.tree_flatten(b)
    if ta != tb:
        return False
    for x, y in zip(la, lb):
        x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
        y = y if isinstance(y, torch.Tensor) else torch.as_tensor(y)
        if x.dtype != y.dtype or x.shape != y.shape:
            return False
        if not torch.allclose(x.cpu(), y.cpu(), rtol=rtol, atol=atol):
            return False
    return True


def construct_matrix(rng, tags, size=3, cond=1000.0):
    """tests/helpers.py:30-83 (rejection-sample cond < 1000)."""
    lx = _lx()
    tags = tags if isinstance(tags, tuple) else (tags,)
    while True:
        m = rng.standard_normal((size, size))
        if lx.diagonal_tag in tags:
            m = np.diag(np.diag(m))
        if lx.symmetric_tag in tags:
            m = m + m.T
        if lx.lower_triangular_tag in tags:
            m = np.tril(m)
        if lx.upper_triangular_tag in tags:
            m = np.triu(m)
        if lx.unit_diagonal_tag in tags:
            m[np.arange(size), np.arange(size)] = 1
        if lx.tridiagonal_tag in tags:
            m = np.diag(np.diag(m)) + np.diag(np.diag(m, 1), 1) + np.diag(np.diag(m, -1), -1)
        if lx.positive_semidefinite_tag in tags:
            m = m @ m.T
        if lx.negative_semidefinite_tag in tags:
            m = -m @ m.T
        if np.linalg.cond(m) < cond:
            return m


def solvers_tags():
    lx = _lx()
    return [
        (lx.AutoLinearSolver(well_posed=True), ()),
        (lx.Triangular(), lx.lower_triangular_tag),
        (lx.Triangular(), lx.upper_triangular_tag),
        (lx.Triangular(), (lx.lower_triangular_tag, lx.unit_diagonal_tag)),
        (lx.Triangular(), (lx.upper_triangular_tag, lx.unit_diagonal_tag)),
        (lx.Diagonal(), lx.diagonal_tag),
        (lx.Diagonal(), (lx.diagonal_tag, lx.unit_diagonal_tag)),
        (lx.Tridiagonal(), lx.tridiagonal_tag),
        (lx.LU(), ()),
        (lx.QR(), ()),
        (lx.BiCGStab(rtol=tol, atol=tol), ()),
        (lx.GMRES(rtol=tol, atol=tol), ()),
        (lx.CG(rtol=tol, atol=tol), lx.positive_semidefinite_tag),
        (lx.CG(rtol=tol, atol=tol), lx.negative_semidefinite_tag),
        (lx.Normal(lx.CG(rtol=tol, atol=tol)), ()),
        (lx.LSMR(atol=tol, rtol=tol), ()),
        (lx.Cholesky(), lx.positive_semidefinite_tag),
        (lx.Cholesky(), lx.negative_semidefinite_tag),
        (lx.Normal(lx.Cholesky()), ()),
    ]


@pytest.mark.parametrize("idx", range(19))
@pytest.mark.parametrize("transpose", [False, True])
def test_small_wellposed(idx, transpose):
    """tests/test_well_posed.py:31-52."""
    lx = _lx()
    solver, tags = solvers_tags()[idx]
    rng = np.random.default_rng(idx)
    cond = math.sqrt(1000) if isinstance(solver, lx.Normal) else 1000
    m = construct_matrix(rng, tags, cond=cond)
    op = lx.MatrixLinearOperator(t64(m), tags)
    if transpose:
        op, m = op.T, m.T
    assert lx.is_symmetric(op) or True
    true_x = rng.standard_normal(3)
    b = t64(m @ true_x)
    sol = lx.linear_solve(op, b, solver=solver, throw=False)
    assert int(sol.result) == 0, lx.RESULTS.name(sol.result)
    np_x = np.linalg.solve(m, b.cpu().numpy())
    assert np.allclose(sol.value.cpu().numpy(), np_x, atol=1e-9, rtol=1e-9)
    assert np.allclose(sol.value.cpu().numpy(), true_x, atol=1e-9, rtol=1e-9)


def test_pytree_wellposed():
    """tests/test_well_posed.py:55-86 (layout of ravel / unravel)."""
    lx = _lx()
    rng = np.random.default_rng(0)
    for solver in (lx.LU(), lx.QR(), lx.GMRES(tol, tol), lx.BiCGStab(tol, tol)):
        while True:
            a, b_, c, d = (rng.standard_normal((3, 3)) for _ in range(4))
            full = np.block([[a, b_], [c, d]])
            if np.linalg.cond(full) < 1000:
                break
        pytree = {"p": {"p": t64(a), "q": t64(b_)}, "q": {"p": t64(c), "q": t64(d)}}
        out_struct = {"p": lx.ShapeDtypeStruct((3,), torch.float64), "q": lx.ShapeDtypeStruct((3,), torch.float64)}
        op = lx.PyTreeLinearOperator(pytree, out_struct)
        tx = {"p": t64(rng.standard_normal(3)), "q": t64(rng.standard_normal(3))}
        bvec = op.mv(tx)
        x = lx.linear_solve(op, bvec, solver, throw=False).value
        assert allclose(x, tx, atol=1e-9, rtol=1e-9)


def test_nontrivial_pytree_operator():
    """tests/test_solve.py:41-48."""
    lx = _lx()
    x = [[1, 5.0], [torch.tensor(-2), torch.tensor(-2.0)]]
    y = [3, 4]
    struct = [lx.ShapeDtypeStruct((), torch.float32)] * 2
    op = lx.PyTreeLinearOperator(x, struct)
    out = lx.linear_solve(op, y).value
    assert allclose(out, [torch.tensor(-3.25), torch.tensor(1.25)])


def test_nontrivial_diagonal_operator():
    """tests/test_solve.py:51-61."""
    lx = _lx()
    x = (8.0, torch.tensor([1, 2, 3]), {"a": torch.tensor([4, 5]), "b": 6})
    y = (4.0, torch.tensor([7, 8, 9]), {"a": torch.tensor([2, 10]), "b": 12})
    out = lx.linear_solve(lx.DiagonalLinearOperator(x), y).value
    true = (torch.tensor(0.5), torch.tensor([7.0, 4.0, 3.0]), {"a": torch.tensor([0.5, 2.0]), "b": torch.tensor(2.0)})
    assert allclose(out, true)


@pytest.mark.parametrize("which", ["LU", "QR"])
def test_mixed_dtypes(which):
    """tests/test_solve.py:64-74: mixed f32/f64 PyTree computes in f64, casts back per leaf."""
    lx = _lx()
    solver = getattr(lx, which)()
    f32 = lambda v: torch.tensor(v, dtype=torch.float32).cuda()
    f64 = lambda v: torch.tensor(v, dtype=torch.float64).cuda()
    x = [[f32(1), f64(5)], [f32(-2), f64(-2)]]
    y = [f64(3), f64(4)]
    struct = [lx.ShapeDtypeStruct((), torch.float64)] * 2
    out = lx.linear_solve(lx.PyTreeLinearOperator(x, struct), y, solver=solver).value
    assert allclose(out, [f32(-3.25), f64(1.25)])


def test_mixed_dtypes_triangular():
    """tests/test_solve.py:103-112."""
    lx = _lx()
    f32 = lambda v: torch.tensor(v, dtype=torch.float32).cuda()
    f64 = lambda v: torch.tensor(v, dtype=torch.float64).cuda()
    x = [[f32(1), f64(0)], [f32(-2), f64(-2)]]
    y = [f64(3), f64(4)]
    struct = [lx.ShapeDtypeStruct((), torch.float64)] * 2
    op = lx.PyTreeLinearOperator(x, struct, lx.lower_triangular_tag)
    out = lx.linear_solve(op, y, solver=lx.Triangular()).value
    assert allclose(out, [f32(3), f64(-5)])


def test_nonfinite_input():
    """tests/test_solve.py:249-261."""
    lx = _lx()
    op = lx.DiagonalLinearOperator((1.0, 1.0))
    for vec in ((1.0, math.inf), (1.0, math.nan), (math.nan, math.inf)):
        sol = lx.linear_solve(op, vec, throw=False)
        assert int(sol.result) == lx.RESULTS.nonfinite_input


def test_iterative_solver_max_steps_only():
    """tests/test_solve.py:174-194."""
    lx = _lx()
    n = 100
    p = -2 * np.eye(n) + np.eye(n, k=1) + np.eye(n, k=-1)
    op = lx.MatrixLinearOperator(t64(p), tags=(lx.negative_semidefinite_tag, lx.symmetric_tag))
    rhs = t64(np.random.default_rng(0).standard_normal(n))
    for solver in (lx.CG(0.0, 0.0, max_steps=2), lx.Normal(lx.CG(0.0, 0.0, max_steps=2)),
                   lx.BiCGStab(0.0, 0.0, max_steps=2), lx.GMRES(0.0, 0.0, max_steps=2),
                   lx.LSMR(0.0, 0.0, max_steps=2)):
        lx.linear_solve(op, rhs, solver)  # must not raise: result is `successful`


def test_singular_iterative_raises():
    """tests/test_singular.py:244-270: iterative solvers on a singular 3x3 raise with throw=True."""
    lx = _lx()
    rng = np.random.default_rng(1)
    m = rng.standard_normal((3, 3))
    m[0, :] = 0
    b = t64(rng.standard_normal(3))
    for solver, tags in ((lx.BiCGStab(tol, tol), ()), (lx.GMRES(tol, tol), ())):
        with pytest.raises(lx.LinearSolveError):
            lx.linear_solve(lx.MatrixLinearOperator(t64(m), tags), b, solver)
    with pytest.raises(lx.LinearSolveError):
        lx.linear_solve(lx.MatrixLinearOperator(t64(m)), b, lx.LU())
    sol = lx.linear_solve(lx.MatrixLinearOperator(t64(m)), b, lx.LU(), throw=False)
    assert int(sol.result) == lx.RESULTS.singular and "non-finite" in lx.RESULTS[sol.result]


def test_nonsquare_qr_lsmr():
    """tests/test_singular.py:102-241: tall and wide least squares vs lstsq."""
    lx = _lx()
    rng = np.random.default_rng(2)
    for shape in ((5, 3), (3, 5), (2, 3), (3, 2)):
        m = rng.standard_normal(shape)
        b = rng.standard_normal(shape[0])
        ref = np.linalg.lstsq(m, b, rcond=None)[0]
        for solver in (lx.QR(), lx.LSMR(tol, tol), lx.AutoLinearSolver(well_posed=None)):
            sol = lx.linear_solve(lx.MatrixLinearOperator(t64(m)), t64(b), solver, throw=False)
            assert np.allclose(sol.value.cpu().numpy(), ref, atol=1e-8), (shape, solver)


def test_lsmr_stats_and_conlim_message():
    """tests/test_lsmr.py:7-30."""
    lx = _lx()
    solver = lx.LSMR(1e-10, 1e-10)
    ill = lx.DiagonalLinearOperator(t64([1e8, 1e6, 1e4, 1e2, 1]))
    try:  # the reference's test only checks the message if the solve raises
        lx.linear_solve(ill, t64(np.ones(5)), solver=solver)
    except lx.LinearSolveError as e:
        assert "Condition number" in str(e)
    with pytest.raises(lx.LinearSolveError, match="Condition number"):
        lx.linear_solve(ill, t64(np.ones(5)), solver=lx.LSMR(1e-10, 1e-10, conlim=1e3))
    sol = lx.linear_solve(ill, t64(np.zeros(5)), solver=solver)
    assert bool((sol.value == 0).all())
    sing = lx.DiagonalLinearOperator(t64([0.0, 4.0, 5.0, 8.0, 10.0]))
    sol = lx.linear_solve(sing, t64([1.0, 0, 0, 0, 0]), solver=solver)
    assert bool((sol.value == 0).all())
    assert set(sol.stats) >= {"num_steps", "istop", "norm_r", "norm_Ar", "norm_A", "cond_A", "norm_x"}


def test_vmap_variants():
    """tests/test_vmap.py:32-89: vmap over operator, vector, both (batch 10) vs lstsq."""
    lx = _lx()
    rng = np.random.default_rng(3)
    for solver, tags in ((lx.LU(), ()), (lx.QR(), ()), (lx.Cholesky(), lx.positive_semidefinite_tag),
                         (lx.CG(tol, tol), lx.positive_semidefinite_tag), (lx.GMRES(tol, tol), ()),
                         (lx.Tridiagonal(), lx.tridiagonal_tag), (lx.LSMR(tol, tol), ())):
        mats = np.stack([construct_matrix(rng, tags) for _ in range(10)])
        vecs = rng.standard_normal((10, 3))
        M, V = t64(mats), t64(vecs)
        f = lambda m, v: lx.linear_solve(lx.MatrixLinearOperator(m, tags), v, solver, throw=False).value
        both = torch.func.vmap(f)(M, V).cpu().numpy()
        ref = np.stack([np.linalg.solve(mats[i], vecs[i]) for i in range(10)])
        assert np.allclose(both, ref, atol=1e-8, rtol=1e-8), solver
        only_m = torch.func.vmap(f, in_dims=(0, None))(M, V[0]).cpu().numpy()
        assert np.allclose(only_m, np.stack([np.linalg.solve(mats[i], vecs[0]) for i in range(10)]), atol=1e-8, rtol=1e-8)
        only_v = torch.func.vmap(f, in_dims=(None, 0))(M[0], V).cpu().numpy()
        assert np.allclose(only_v, np.stack([np.linalg.solve(mats[0], vecs[i]) for i in range(10)]), atol=1e-8, rtol=1e-8)


def test_vmap_vmap():
    """tests/test_vmap_vmap.py: nested vmap, inner over vectors, outer over operators."""
    lx = _lx()
    rng = np.random.default_rng(4)
    mats = np.stack([construct_matrix(rng, ()) for _ in range(4)])
    vecs = rng.standard_normal((4, 5, 3))
    f = lambda m, v: lx.linear_solve(lx.MatrixLinearOperator(m), v, lx.LU(), throw=False).value
    out = torch.func.vmap(torch.func.vmap(f, in_dims=(None, 0)))(t64(mats), t64(vecs)).cpu().numpy()
    ref = np.stack([[np.linalg.solve(mats[i], vecs[i, j]) for j in range(5)] for i in range(4)])
    assert np.allclose(out, ref, atol=1e-9)
    res = torch.func.vmap(lambda m, v: lx.linear_solve(lx.MatrixLinearOperator(m), v, lx.LU(), throw=False).result)(
        t64(mats), t64(vecs[:, 0]))
    assert res.shape == (4,) and int(res.abs().sum()) == 0


def test_state_reuse_and_transpose():
    """_solve.py:732-740 (state=) and solver.transpose == init(operator.T) (test_transpose.py)."""
    lx = _lx()
    rng = np.random.default_rng(5)
    m = construct_matrix(rng, (), size=6)
    op = lx.MatrixLinearOperator(t64(m))
    b = t64(rng.standard_normal(6))
    for solver in (lx.LU(), lx.QR()):
        state = solver.init(op, {})
        x1 = lx.linear_solve(op, b, solver, state=state).value.cpu().numpy()
        assert np.allclose(x1, np.linalg.solve(m, b.cpu().numpy()), atol=1e-10)
        t_state, t_opts = solver.transpose(state, {})
        x2 = lx.linear_solve(op.T, b, solver, state=t_state, options=t_opts).value.cpu().numpy()
        assert np.allclose(x2, np.linalg.solve(m.T, b.cpu().numpy()), atol=1e-10)
    sol = lx.linear_solve(op, b, lx.LU())
    (lu, piv), _, transposed = sol.state  # lazily materialised state
    assert lu.shape == (6, 6) and piv.dtype == torch.int32 and transposed is False


def test_operator_mv_and_norms():
    lx = _lx()
    rng = np.random.default_rng(6)
    m, v = rng.standard_normal((7, 5)), rng.standard_normal(5)
    assert np.allclose(lx.MatrixLinearOperator(t64(m)).mv(t64(v)).cpu().numpy(), m @ v)
    assert np.allclose(lx.MatrixLinearOperator(t64(m)).T.mv(t64(m @ v)).cpu().numpy(), m.T @ (m @ v))
    d, l, u = rng.standard_normal(6), rng.standard_normal(5), rng.standard_normal(5)
    T = np.diag(d) + np.diag(l, -1) + np.diag(u, 1)
    w = rng.standard_normal(6)
    top = lx.TridiagonalLinearOperator(t64(d), t64(l), t64(u))
    assert np.allclose(top.mv(t64(w)).cpu().numpy(), T @ w)
    assert np.allclose(top.as_matrix().cpu().numpy(), T)
    x = [t64(rng.standard_normal(4)), t64(rng.standard_normal((2, 3)))]
    flat = np.concatenate([x[0].cpu().numpy().ravel(), x[1].cpu().numpy().ravel()])
    assert np.isclose(float(lx.two_norm(x)), np.linalg.norm(flat))
    assert np.isclose(float(lx.max_norm(x)), np.abs(flat).max())
    assert np.isclose(float(lx.tree_dot(x, x)), flat @ flat)
    assert float(lx.two_norm(t64([-3.0]))) == 3.0


def test_gram_kernel_matches_float64():
    """lxb_gram_* (the operator of the normal equations, normal.py:111-117)."""
    from lineax_b200 import _ops

    rng = np.random.default_rng(0)
    for (m, n), dt in (((5, 3), np.float64), ((3, 5), np.float64), ((130, 70), np.float32), ((64, 200), np.float32)):
        a = rng.standard_normal((2, m, n)).astype(dt)
        A = torch.as_tensor(a).cuda()
        g1 = _ops.gram(A, False).cpu().numpy()
        g2 = _ops.gram(A, True).cpu().numpy()
        a64 = a.astype(np.float64)
        tol = 1e-12 if dt == np.float64 else 1e-5
        assert np.max(np.abs(g1 - np.einsum("bki,bkj->bij", a64, a64))) <= tol * np.max(np.abs(g1))
        assert np.max(np.abs(g2 - np.einsum("bik,bjk->bij", a64, a64))) <= tol * np.max(np.abs(g2))
        assert np.array_equal(g1, np.swapaxes(g1, -1, -2)), "exactly symmetric"


@pytest.mark.parametrize("shape", [(12, 7), (7, 12)])
def test_normal_cg_preconditioner_and_y0(shape):
    """lineax/_solver/normal.py:31-52: an outer preconditioner M ~ pinv(A) reaches the inner CG as
    M M^* (tall) / M^* M (wide), tagged positive semidefinite; y0 is mapped to M^* y0 in the wide case.
    With the exact pseudo-inverse the preconditioned normal equations converge in a handful of steps,
    and the solution must be the least-squares / minimum-norm one."""
    import oracle

    lx = _lx()
    m, n = shape
    rng = np.random.default_rng(m * 31 + n)
    a = rng.standard_normal((m, n))
    b = rng.standard_normal(m)
    pinv = np.linalg.pinv(a)
    x_ref = np.linalg.lstsq(a, b, rcond=None)[0]
    op = lx.MatrixLinearOperator(t64(a))
    solver = lx.Normal(lx.CG(rtol=1e-10, atol=1e-10))
    plain = lx.linear_solve(op, t64(b), solver)
    pre = lx.linear_solve(op, t64(b), solver, options={"preconditioner": lx.MatrixLinearOperator(t64(pinv))})
    assert np.allclose(plain.value.cpu().numpy(), x_ref, rtol=1e-7, atol=1e-9)
    assert np.allclose(pre.value.cpu().numpy(), x_ref, rtol=1e-7, atol=1e-9)
    assert int(pre.stats["num_steps"]) <= 4 <= int(plain.stats["num_steps"]) + 4
    # oracle parity of the inner solve: CG on the Gram matrix with the squared preconditioner
    tall = m >= n
    gram = a.T @ a if tall else a @ a.T
    msq = pinv @ pinv.T if tall else pinv.T @ pinv
    rhs = a.T @ b if tall else b
    yr, rr, st = oracle.cg(gram, rhs, 1e-10, 1e-10, preconditioner=msq)
    xr = yr if tall else a.T @ yr
    assert int(pre.result) == rr == 0 and abs(int(pre.stats["num_steps"]) - st["num_steps"]) <= 2
    assert np.allclose(pre.value.cpu().numpy(), xr, rtol=1e-8, atol=1e-10)
    # y0: the exact solution as initial guess stops immediately (tall: y0 is passed through)
    if tall:
        warm = lx.linear_solve(op, t64(b), solver, options={"y0": t64(x_ref)})
        assert int(warm.stats["num_steps"]) <= 1
    assert solver.assume_full_rank() is True
    assert lx.Normal(lx.Cholesky()).assume_full_rank() == lx.Cholesky().assume_full_rank()
