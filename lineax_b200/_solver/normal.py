"""`Normal`, mirroring lineax/_solver/normal.py:31-181: solve A^T A x = A^T b (tall) or
A A^T y = b, x = A^T y (wide) with an inner solver for positive definite systems."""
from __future__ import annotations

from .. import _ops
from .. import _tree as tr
from .._operator import MatrixLinearOperator, TaggedLinearOperator
from .._solve import AbstractLinearSolver
from .._tags import positive_semidefinite_tag
from .misc import ravel_leaves, unravel_like


class Normal(AbstractLinearSolver):
    """Wraps an inner solver (e.g. `CG`, `Cholesky`) to solve via the normal equations.

    For materialised operators the Gram matrix is built once in `init` with the native GEMV
    kernel (the columns of A form the batch).  state = (inner_state, operator, tall, gram_op)."""

    def __init__(self, inner_solver: AbstractLinearSolver):
        self.inner_solver = inner_solver

    def init(self, operator, options):
        a = operator.as_matrix()
        m, n = a.shape
        tall = m >= n
        gram = _gram(a, tall)
        gop = MatrixLinearOperator(gram, positive_semidefinite_tag)
        inner_options = {k: v for k, v in options.items() if k not in ("y0",)}
        if "preconditioner" in inner_options and not tall:
            inner_options.pop("preconditioner")
        return self.inner_solver.init(gop, inner_options), operator, tall, gop

    def compute(self, state, vector, options):
        inner_state, operator, tall, gop = state
        a = operator.as_matrix()
        b = ravel_leaves(tr.tree_leaves(vector))
        inner_options = {k: v for k, v in options.items() if k != "y0"}
        if tall:
            rhs = _ops.matvec(a, b, True)  # A^T b
            if "y0" in options:
                inner_options["y0"] = ravel_leaves(tr.tree_leaves(options["y0"]))
            sol, result, stats = self.inner_solver.compute(inner_state, rhs, inner_options)
        else:
            inner_options.pop("preconditioner", None)
            y, result, stats = self.inner_solver.compute(inner_state, b, inner_options)
            sol = _ops.matvec(a, y, True)  # x = A^T y
        return unravel_like(sol, operator.in_structure()), result, stats

    def transpose(self, state, options):
        inner_state, operator, tall, gop = state
        new = self.init(operator.transpose(), options)
        return new, options

    def conj(self, state, options):
        return state, options

    def assume_full_rank(self):
        return True


def _gram(a, tall):
    """G[i, :] = M^T (M e_i-th column) computed with the native GEMV, one batched launch:
    treat the n columns of M as a batch of vectors."""
    m_ = (a if tall else a.mT).contiguous()  # G = m_^T m_
    cols = m_.mT.contiguous()  # [k, rows]: row i is column i of m_
    # G[i, j] = <col_i, col_j> = (m_^T col_i)[j]: batched transposed matvec, matrix broadcast
    return _ops.matvec(m_, cols, True)
