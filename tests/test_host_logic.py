"""Host-side logic on CPU (no GPU): operator structure queries, AutoLinearSolver dispatch,
error behaviour, PyTree packing, vmap batching rules and result rewriting.  The numerical ops are
replaced by oracle-backed CPU test doubles (tests/cpu_kernels.py); the product itself has no CPU path."""
import math

import numpy as np
import pytest
import torch

import lineax_b200 as lx
from tests import cpu_kernels

t64 = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64))


@pytest.fixture(autouse=True, scope="module")
def _cpu_doubles():
    """Install the CPU test doubles and route host-side scalars to the CPU for this module only."""
    from lineax_b200 import _tree

    cpu_kernels.install()
    old = _tree._default_device
    lx.set_default_device("cpu")
    yield
    _tree._default_device = old


def test_product_has_no_cpu_fallback():
    """Without the test doubles an op on CPU tensors must fail loudly."""
    with pytest.raises((NotImplementedError, RuntimeError)):
        lx._ops.cholesky_factor(torch.eye(3, dtype=torch.float64), False)


def test_tags_and_queries():
    m = torch.randn(3, 3, dtype=torch.float64)
    op = lx.MatrixLinearOperator(m, (lx.lower_triangular_tag, lx.unit_diagonal_tag))
    assert lx.is_lower_triangular(op) and not lx.is_upper_triangular(op) and lx.has_unit_diagonal(op)
    assert lx.is_upper_triangular(op.T) and not lx.is_lower_triangular(op.T)
    psd = lx.MatrixLinearOperator(m, lx.positive_semidefinite_tag)
    assert lx.is_symmetric(psd) and lx.is_positive_semidefinite(psd) and psd.transpose() is psd
    assert lx.is_negative_semidefinite(-psd) and not lx.is_positive_semidefinite(-psd)
    tagged = lx.TaggedLinearOperator(lx.MatrixLinearOperator(m), lx.tridiagonal_tag)
    assert lx.is_tridiagonal(tagged) and not lx.is_diagonal(tagged)
    d = lx.DiagonalLinearOperator(torch.ones(4))
    assert lx.is_diagonal(d) and lx.is_symmetric(d) and lx.is_tridiagonal(d) and d.T is d
    assert lx.transpose_tags(frozenset([lx.lower_triangular_tag])) == frozenset([lx.upper_triangular_tag])
    assert lx.is_diagonal(lx.MatrixLinearOperator(torch.ones(1, 1)))


def test_auto_solver_dispatch_table():
    """lineax/_solve.py:555-600."""
    sq = torch.randn(3, 3)
    pick = lambda wp, op: type(lx.AutoLinearSolver(wp).select_solver(op)).__name__
    assert pick(True, lx.MatrixLinearOperator(sq)) == "LU"
    assert pick(None, lx.MatrixLinearOperator(torch.randn(3, 4))) == "QR"
    assert pick(True, lx.MatrixLinearOperator(sq, lx.positive_semidefinite_tag)) == "Cholesky"
    assert pick(True, lx.MatrixLinearOperator(sq, lx.negative_semidefinite_tag)) == "Cholesky"
    assert pick(True, lx.MatrixLinearOperator(sq, lx.upper_triangular_tag)) == "Triangular"
    assert pick(True, lx.MatrixLinearOperator(sq, lx.tridiagonal_tag)) == "Tridiagonal"
    assert pick(True, lx.DiagonalLinearOperator(torch.ones(3))) == "Diagonal"
    assert lx.AutoLinearSolver(True).select_solver(lx.DiagonalLinearOperator(torch.ones(3))).well_posed
    assert not lx.AutoLinearSolver(False).select_solver(lx.DiagonalLinearOperator(torch.ones(3))).well_posed
    with pytest.raises(ValueError, match="non-square"):
        lx.AutoLinearSolver(True).select_solver(lx.MatrixLinearOperator(torch.randn(3, 4)))
    with pytest.raises(NotImplementedError, match="SVD"):
        lx.AutoLinearSolver(False).select_solver(lx.MatrixLinearOperator(sq))
    with pytest.raises(ValueError, match="Invalid value"):
        lx.AutoLinearSolver("yes").select_solver(lx.MatrixLinearOperator(sq))


def test_argument_errors_match_reference():
    sq = lx.MatrixLinearOperator(torch.randn(3, 3))
    with pytest.raises(ValueError, match="should be an `AbstractLinearOperator`"):
        lx.linear_solve(torch.randn(3, 3), torch.randn(3))
    with pytest.raises(ValueError, match="structures do not match"):
        lx.linear_solve(sq, torch.randn(4))
    with pytest.raises(ValueError, match="2-dimensional"):
        lx.MatrixLinearOperator(torch.randn(3))
    with pytest.raises(ValueError, match="positive or negative definite"):
        lx.CG(1e-6, 1e-6).init(sq, {})
    with pytest.raises(ValueError, match="square"):
        lx.LU().init(lx.MatrixLinearOperator(torch.randn(3, 4)), {})
    with pytest.raises(ValueError, match="tridiagonal"):
        lx.Tridiagonal().init(sq, {})
    with pytest.raises(ValueError, match="non-negative"):
        lx.CG(-1.0, 1e-6)
    with pytest.raises(ValueError, match="Must specify"):
        lx.GMRES(0.0, 0.0)
    with pytest.raises(ValueError, match="consistent size"):
        lx.TridiagonalLinearOperator(torch.ones(3), torch.ones(3), torch.ones(2))
    with pytest.raises(ValueError, match="preconditioner must be a linear operator"):
        lx.linear_solve(lx.MatrixLinearOperator(torch.eye(3), lx.positive_semidefinite_tag), torch.ones(3),
                        lx.CG(1e-6, 1e-6), options={"preconditioner": torch.eye(3)})


def test_results_enumeration():
    assert [lx.RESULTS.successful, lx.RESULTS.max_steps_reached, lx.RESULTS.singular, lx.RESULTS.breakdown,
            lx.RESULTS.stagnation, lx.RESULTS.conlim, lx.RESULTS.nonfinite_input] == list(range(7))
    assert lx.RESULTS[0] == "" and "maximum number of solver steps" in lx.RESULTS[torch.tensor(1)]
    assert "Condition number" in lx.RESULTS[5] and "non-finite" in lx.RESULTS[6]


def test_pytree_packing_and_kat():
    """tests/test_solve.py:41-61 through the full host path (PyTree flatten, promote, unravel)."""
    x = [[1, 5.0], [torch.tensor(-2), torch.tensor(-2.0)]]
    struct = [lx.ShapeDtypeStruct((), torch.float32)] * 2
    op = lx.PyTreeLinearOperator(x, struct)
    assert torch.equal(op.as_matrix(), torch.tensor([[1.0, 5.0], [-2.0, -2.0]]))
    assert torch.equal(op.T.as_matrix(), op.as_matrix().T)
    out = lx.linear_solve(op, [3, 4]).value
    assert torch.allclose(torch.stack(out), torch.tensor([-3.25, 1.25]))
    d = (8.0, torch.tensor([1, 2, 3]), {"a": torch.tensor([4, 5]), "b": 6})
    y = (4.0, torch.tensor([7, 8, 9]), {"a": torch.tensor([2, 10]), "b": 12})
    out = lx.linear_solve(lx.DiagonalLinearOperator(d), y).value
    assert torch.allclose(out[1], torch.tensor([7.0, 4.0, 3.0])) and float(out[2]["b"]) == 2.0
    # block operator with array leaves: row blocks = out leaves, column blocks = in leaves
    a, b, c, e = (torch.randn(2, 2, dtype=torch.float64) for _ in range(4))
    op = lx.PyTreeLinearOperator({"u": {"u": a, "v": b}, "v": {"u": c, "v": e}},
                                 {"u": lx.ShapeDtypeStruct((2,), torch.float64), "v": lx.ShapeDtypeStruct((2,), torch.float64)})
    assert torch.equal(op.as_matrix(), torch.cat([torch.cat([a, b], 1), torch.cat([c, e], 1)], 0))


def test_nonfinite_and_throw():
    op = lx.DiagonalLinearOperator((1.0, 1.0))
    sol = lx.linear_solve(op, (1.0, math.inf), throw=False)
    assert int(sol.result) == lx.RESULTS.nonfinite_input
    with pytest.raises(lx.LinearSolveError, match="non-finite"):
        lx.linear_solve(op, (1.0, math.nan))
    sing = lx.MatrixLinearOperator(t64(np.zeros((3, 3))))
    sol = lx.linear_solve(sing, t64(np.ones(3)), lx.LU(), throw=False)
    assert int(sol.result) == lx.RESULTS.singular
    ident = lx.IdentityLinearOperator(lx.ShapeDtypeStruct((3,), torch.float64))
    v = t64([1.0, 2.0, 3.0])
    assert lx.linear_solve(ident, v).value is v  # _solve.py:778-784 short-circuit


def test_vmap_batching_rules_cpu():
    """torch.func.vmap over operator / vector / both / nested consumes batch dims natively."""
    rng = np.random.default_rng(0)
    mats = rng.standard_normal((6, 4, 4)) + 4 * np.eye(4)
    vecs = rng.standard_normal((6, 4))
    M, V = t64(mats), t64(vecs)
    f = lambda m, v: lx.linear_solve(lx.MatrixLinearOperator(m), v, lx.LU(), throw=False).value
    ref = np.stack([np.linalg.solve(mats[i], vecs[i]) for i in range(6)])
    assert np.allclose(torch.func.vmap(f)(M, V).numpy(), ref)
    assert np.allclose(torch.func.vmap(f, in_dims=(0, None))(M, V[0]).numpy(),
                       np.stack([np.linalg.solve(mats[i], vecs[0]) for i in range(6)]))
    assert np.allclose(torch.func.vmap(f, in_dims=(None, 0))(M[0], V).numpy(),
                       np.stack([np.linalg.solve(mats[0], vecs[i]) for i in range(6)]))
    nested = torch.func.vmap(torch.func.vmap(f, in_dims=(None, 0)))(M[:3], V.reshape(3, 2, 4))
    for i in range(3):
        for j in range(2):
            assert np.allclose(nested[i, j].numpy(), np.linalg.solve(mats[i], vecs[2 * i + j]))
    g = lambda m, v: lx.linear_solve(lx.MatrixLinearOperator(m, lx.positive_semidefinite_tag), v,
                                     lx.CG(1e-10, 1e-10), throw=False)
    spd = np.einsum("bij,bkj->bik", mats, mats)
    sol = torch.func.vmap(lambda m, v: (g(m, v).value, g(m, v).result, g(m, v).stats["num_steps"]))(t64(spd), V)
    assert sol[1].shape == (6,) and int(sol[1].abs().sum()) == 0 and sol[2].shape == (6,)
    assert np.allclose(sol[0].numpy(), np.stack([np.linalg.solve(spd[i], vecs[i]) for i in range(6)]), atol=1e-7)
    # throw=True inside vmap reports how many systems failed
    bad = M.clone()
    bad[2] = 0
    with pytest.raises(lx.LinearSolveError, match="1 of 6"):
        torch.func.vmap(lambda m, v: lx.linear_solve(lx.MatrixLinearOperator(m), v, lx.LU()).value)(bad, V)


def test_state_transpose_contract():
    """`transpose(state)` equals `init(operator.T)` (lineax/_solve.py:412-419) for LU / QR."""
    rng = np.random.default_rng(1)
    m = rng.standard_normal((5, 5)) + 3 * np.eye(5)
    op = lx.MatrixLinearOperator(t64(m))
    b = t64(rng.standard_normal(5))
    for solver in (lx.LU(), lx.QR()):
        st = solver.init(op, {})
        tst, topt = solver.transpose(st, {})
        x = lx.linear_solve(op.T, b, solver, state=tst, options=topt).value.numpy()
        assert np.allclose(x, np.linalg.solve(m.T, b.numpy()))
    sol = lx.linear_solve(op, b, lx.LU())
    (lu, piv), packed, transposed = sol.state
    assert piv.dtype == torch.int32 and lu.shape == (5, 5) and transposed is False
    assert lx.LU().assume_full_rank() and not lx.LSMR(1e-6, 1e-6).assume_full_rank()
    assert lx.LU() == lx.LU() and lx.CG(1e-6, 1e-6) != lx.CG(1e-5, 1e-6)
