"""`Normal`, mirroring lineax/_solver/normal.py:31-181: solve A^T A x = A^T b (tall) or
A A^T y = b, x = A^T y (wide) with an inner solver for positive (semi)definite systems."""
from __future__ import annotations

from copy import copy

from .. import _ops
from .. import _tree as tr
from .._operator import MatrixLinearOperator, conj, linearise
from .._solve import AbstractLinearSolver
from .._tags import positive_semidefinite_tag
from .misc import ravel_leaves, unravel_like


def normal_preconditioner_and_y0(options: dict, tall: bool) -> dict:
    """lineax/_solver/normal.py:31-52: the outer preconditioner M (an approximation of the pseudo-inverse
    of A) becomes M M^* (tall) or M^* M (wide), tagged positive semidefinite, for the inner solver; in the
    wide case an initial guess y0 is mapped to M^* y0."""
    preconditioner = options.get("preconditioner")
    y0 = options.get("y0")
    inner_options = copy(options)
    if preconditioner is not None:
        pm = linearise(preconditioner).as_matrix()
        if tall:
            inner_options["preconditioner"] = MatrixLinearOperator(_ops.gram(pm, True), positive_semidefinite_tag)
        else:
            inner_options["preconditioner"] = MatrixLinearOperator(_ops.gram(pm, False), positive_semidefinite_tag)
            if y0 is not None:
                y0f = ravel_leaves(tr.tree_leaves(y0))
                inner_options["y0"] = _ops.matvec(pm, y0f.to(pm.dtype), True)
    return inner_options


class Normal(AbstractLinearSolver):
    """Wraps an inner solver (e.g. `CG`, `Cholesky`) to solve via the normal equations.

    For materialised operators the Gram matrix is built once in `init` by the native tiled kernel
    (`lxb_gram_*`).  state = (inner_state, tall, operator, inner_options) -- normal.py:107-119."""

    def __init__(self, inner_solver: AbstractLinearSolver):
        self.inner_solver = inner_solver

    def init(self, operator, options):
        tall = operator.out_size() >= operator.in_size()
        a = linearise(operator).as_matrix()
        gop = MatrixLinearOperator(_ops.gram(a, not tall), positive_semidefinite_tag)
        inner_options = normal_preconditioner_and_y0(options, tall)
        inner_state = self.inner_solver.init(gop, inner_options)
        return inner_state, tall, operator, inner_options

    def compute(self, state, vector, options):
        inner_state, tall, operator, inner_options = state
        del options  # normal.py:130: the options fixed at init are the ones used
        a = operator.as_matrix()
        b = ravel_leaves(tr.tree_leaves(vector))
        if tall:
            b = _ops.matvec(a, b, True)  # A^* b
            if "y0" in inner_options and not hasattr(inner_options["y0"], "shape"):
                inner_options = dict(inner_options, y0=ravel_leaves(tr.tree_leaves(inner_options["y0"])))
        sol, result, stats = self.inner_solver.compute(inner_state, b, inner_options)
        if not tall:
            sol = _ops.matvec(a, sol, True)  # x = A^* y
        return unravel_like(sol, operator.in_structure()), result, stats

    def transpose(self, state, options):
        inner_state, tall, operator, inner_options = state
        inner_state_conj, inner_options = self.inner_solver.conj(inner_state, inner_options)
        return (inner_state_conj, not tall, operator.transpose(), inner_options), options

    def conj(self, state, options):
        inner_state, tall, operator, inner_options = state
        inner_state_conj, inner_options = self.inner_solver.conj(inner_state, inner_options)
        return (inner_state_conj, tall, conj(operator), inner_options), options

    def assume_full_rank(self):
        return self.inner_solver.assume_full_rank()
