"""`Cholesky`, drop-in for lineax/_solver/cholesky.py:34-95 on csrc/direct.cu."""
from __future__ import annotations

from .. import _ops
from .. import _tree as tr
from .._operator import is_negative_semidefinite, is_positive_semidefinite
from .._solution import RESULTS
from .._solve import AbstractLinearSolver
from .misc import ravel_leaves, unravel_like


class Cholesky(AbstractLinearSolver):
    """Cholesky solver (square, nonsingular, positive or negative definite operators).

    state = (factor_upper, is_nsd) with (+-A) = U^T U  -- cholesky.py:31,63.
    """

    def init(self, operator, options):
        del options
        is_nsd = is_negative_semidefinite(operator)
        if not (is_positive_semidefinite(operator) | is_nsd):
            raise ValueError(
                "`Cholesky(..., normal=False)` may only be used for positive "
                "or negative definite linear operators"
            )
        matrix = operator.as_matrix()
        m, n = matrix.shape
        if m != n:
            raise ValueError(
                "`Cholesky(..., normal=False)` may only be used for linear solves with square matrices"
            )
        return _ops.cholesky_factor(matrix, bool(is_nsd)), is_nsd

    def compute(self, state, vector, options):
        factor, is_nsd = state
        del options
        flat = ravel_leaves(tr.tree_leaves(vector))
        solution = _ops.cholesky_solve(factor, flat, bool(is_nsd))
        return unravel_like(solution, tr.struct_of(vector)), RESULTS.successful, {}

    def transpose(self, state, options):
        factor, is_nsd = state  # matrix is self-adjoint
        return (factor.conj(), is_nsd), options

    def conj(self, state, options):
        factor, is_nsd = state
        return (factor.conj(), is_nsd), options

    def assume_full_rank(self):
        return True
