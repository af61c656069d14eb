"""`BiCGStab`, drop-in for lineax/_solver/bicgstab.py:35-222 on csrc/bicgstab.cu."""
from __future__ import annotations

import torch

from .. import _native as nat
from .. import _ops
from .. import _tree as tr
from .._norm import max_norm
from .._operator import conj, linearise
from .._solve import AbstractLinearSolver, config
from ._iterative import check_tols, conj_options, flat_problem, steps_flags, transpose_options
from .misc import unravel_like


class BiCGStab(AbstractLinearSolver):
    """Biconjugate gradient stabilised method (square operators, right preconditioning).

    Options: `preconditioner`, `y0`.  state = operator.
    """

    def __init__(self, rtol, atol, norm=max_norm, max_steps=None):
        self.rtol, self.atol, self.norm, self.max_steps = rtol, atol, norm, max_steps
        check_tols(self)
        if norm is not max_norm:
            raise NotImplementedError("the native BiCGStab kernel implements the default `max_norm` test")

    def init(self, operator, options):
        from .gmres import _is_row_sharded

        if _is_row_sharded(operator):
            if operator.rows != operator.cols:
                raise ValueError(
                    "`BiCGstab(..., normal=False)` may only be used for linear solves with square matrices."
                )
            return operator
        if not tr.structure_equal(operator.in_structure(), operator.out_structure()):
            raise ValueError(
                "`BiCGstab(..., normal=False)` may only be used for linear solves with square matrices."
            )
        return linearise(operator)

    def compute(self, state, vector, options):
        operator = state
        from .gmres import _is_row_sharded

        if _is_row_sharded(operator):
            from ..distributed import RowShardedBiCGStab

            if options.get("preconditioner") is not None:
                raise NotImplementedError("row-sharded BiCGStab takes no preconditioner")
            x64 = config.enable_x64
            if x64 is None:
                x64 = operator.local.dtype == torch.float64
            key = (operator.rows, float(self.rtol), float(self.atol), self.max_steps, bool(x64), operator.local.dtype)
            solver = operator.sharded_solver("bicgstab", key, lambda: RowShardedBiCGStab(
                operator.rows, float(self.rtol), float(self.atol), max_steps=self.max_steps, x64=x64,
                dtype=operator.local.dtype, group=operator.group))
            x, result, steps = solver.solve(operator.local, vector, options.get("y0"))
            return x, result, {"num_steps": steps, "max_steps": self.max_steps}
        a, b, m, y0, size, _ = flat_problem(operator, vector, options)
        ms, flags = steps_flags(self.max_steps, size)
        x64 = config.enable_x64
        if x64 is None:  # bicgstab.py:110 keys on jax_enable_x64; float64 data implies it
            x64 = a.dtype == torch.float64
        if x64:
            flags |= nat.X64_BREAKDOWN
        x, result, steps = _ops.bicgstab(a, b, m, y0, float(self.rtol), float(self.atol), ms, flags)
        return unravel_like(x, tr.struct_of(vector)), result, {"num_steps": steps, "max_steps": self.max_steps}

    def transpose(self, state, options):
        return state.transpose(), transpose_options(options)

    def conj(self, state, options):
        return conj(state), conj_options(options)

    def assume_full_rank(self):
        return True
