"""`GMRES`, drop-in for lineax/_solver/gmres.py:39-430 on csrc/gmres.cu."""
from __future__ import annotations

from .. import _ops
from .. import _tree as tr
from .._norm import max_norm
from .._operator import conj, linearise
from .._solve import AbstractLinearSolver
from ._iterative import check_tols, conj_options, flat_problem, steps_flags, transpose_options
from .misc import unravel_like


class GMRES(AbstractLinearSolver):
    """Restarted GMRES (square operators, left preconditioning).

    `num_steps` counts restarts (outer iterations, including the dummy first pass), exactly
    as in the reference.  Options: `preconditioner`, `y0`.  state = operator.
    """

    def __init__(self, rtol, atol, norm=max_norm, max_steps=None, restart=20, stagnation_iters=20):
        self.rtol, self.atol, self.norm, self.max_steps = rtol, atol, norm, max_steps
        self.restart, self.stagnation_iters = restart, stagnation_iters
        check_tols(self)
        if norm is not max_norm:
            raise NotImplementedError("the native GMRES kernel implements the default `max_norm` test")

    def init(self, operator, options):
        if _is_row_sharded(operator):
            if operator.rows != operator.cols:
                raise ValueError(
                    "`GMRES(..., normal=False)` may only be used for linear solves with square matrices."
                )
            return operator
        del options
        if not tr.structure_equal(operator.in_structure(), operator.out_structure()):
            raise ValueError(
                "`GMRES(..., normal=False)` may only be used for linear solves with square matrices."
            )
        return linearise(operator)

    def compute(self, state, vector, options):
        operator = state
        if _is_row_sharded(operator):
            return self._compute_row_sharded(operator, vector, options)
        a, b, m, y0, size, _ = flat_problem(operator, vector, options)
        ms, flags = steps_flags(self.max_steps, size)
        restart = min(int(self.restart), size)  # gmres.py:128
        x, result, steps = _ops.gmres(a, b, m, y0, float(self.rtol), float(self.atol), ms, restart,
                                      int(self.stagnation_iters), flags)
        return unravel_like(x, tr.struct_of(vector)), result, {"num_steps": steps, "max_steps": self.max_steps}

    def _compute_row_sharded(self, operator, vector, options):
        """ONE system row-partitioned over the GPUs (csrc/gmres_dist.cu): `vector` and the solution are this
        rank's row slices; result / num_steps are identical on every rank."""
        from ..distributed import RowShardedGMRES

        if options.get("preconditioner") is not None:
            raise NotImplementedError("row-sharded GMRES takes no preconditioner")
        n = operator.rows
        key = (n, float(self.rtol), float(self.atol), int(self.restart), int(self.stagnation_iters), self.max_steps,
               operator.local.dtype)
        solver = operator.sharded_solver("gmres", key, lambda: RowShardedGMRES(
            n, float(self.rtol), float(self.atol), restart=int(self.restart),
            stagnation_iters=int(self.stagnation_iters), max_steps=self.max_steps, dtype=operator.local.dtype,
            group=operator.group))
        x, result, steps = solver.solve(operator.local, vector, options.get("y0"))
        return x, result, {"num_steps": steps, "max_steps": self.max_steps}

    def transpose(self, state, options):
        return state.transpose(), transpose_options(options)

    def conj(self, state, options):
        return conj(state), conj_options(options)

    def assume_full_rank(self):
        return True


def _is_row_sharded(operator) -> bool:
    from ..distributed import RowShardedMatrixLinearOperator

    return isinstance(operator, RowShardedMatrixLinearOperator)
