"""`CG` solver, drop-in for lineax/_solver/cg.py:48-250 on the fused persistent kernel (csrc/cg.cu)."""
from __future__ import annotations

import warnings

from .. import _native as nat
from .. import _ops
from .. import _tree as tr
from .._norm import max_norm
from .._operator import (
    conj, is_negative_semidefinite, is_positive_semidefinite, linearise,
)
from .._solve import AbstractLinearSolver
from .misc import preconditioner_and_y0, ravel_leaves, unravel_like


def _check_tols(self):
    if isinstance(self.rtol, (int, float)) and self.rtol < 0:
        raise ValueError("Tolerances must be non-negative.")
    if isinstance(self.atol, (int, float)) and self.atol < 0:
        raise ValueError("Tolerances must be non-negative.")
    if isinstance(self.atol, (int, float)) and isinstance(self.rtol, (int, float)):
        if self.atol == 0 and self.rtol == 0 and self.max_steps is None:
            raise ValueError(
                "Must specify `rtol`, `atol`, or `max_steps` (or some combination of all three)."
            )


class CG(AbstractLinearSolver):
    """Conjugate gradient solver for positive or negative definite operators.

    Options: `preconditioner` (positive definite operator, left preconditioning) and `y0`.
    state = (operator, is_nsd)  -- lineax/_solver/cg.py:42,103.
    """

    def __init__(self, rtol, atol, norm=max_norm, stabilise_every=10, max_steps=None):
        self.rtol, self.atol, self.norm = rtol, atol, norm
        self.stabilise_every, self.max_steps = stabilise_every, max_steps
        _check_tols(self)
        if norm is not max_norm:
            raise NotImplementedError("the native CG kernel implements the default `max_norm` test")

    def init(self, operator, options):
        from .gmres import _is_row_sharded

        if _is_row_sharded(operator):  # tags of a sharded operator are the caller's word (`operator.tags`)
            if operator.rows != operator.cols:
                raise ValueError("`CG()` may only be used for linear solves with square matrices.")
            return operator, bool(getattr(operator, "is_nsd", False))
        del options
        is_nsd = is_negative_semidefinite(operator)
        if not tr.structure_equal(operator.in_structure(), operator.out_structure()):
            raise ValueError("`CG()` may only be used for linear solves with square matrices.")
        if not (is_positive_semidefinite(operator) | is_nsd):
            raise ValueError(
                "`CG()` may only be used for positive or negative definite linear operators"
            )
        # cg.py:100-101 negates the operator here; the kernel applies the sign instead (LXB_NSD)
        return linearise(operator), is_nsd

    def compute(self, state, vector, options):
        operator, is_nsd = state
        from .gmres import _is_row_sharded

        if _is_row_sharded(operator):
            from ..distributed import RowShardedCG

            if options.get("preconditioner") is not None:
                raise NotImplementedError("row-sharded CG takes no preconditioner")
            key = (operator.rows, float(self.rtol), float(self.atol), self.stabilise_every, self.max_steps, is_nsd,
                   operator.local.dtype)
            solver = operator.sharded_solver("cg", key, lambda: RowShardedCG(
                operator.rows, float(self.rtol), float(self.atol), stabilise_every=self.stabilise_every,
                max_steps=self.max_steps, is_nsd=is_nsd, dtype=operator.local.dtype, group=operator.group))
            x, result, steps = solver.solve(operator.local, vector, options.get("y0"))
            return x, result, {"num_steps": steps, "max_steps": self.max_steps}
        preconditioner, y0 = preconditioner_and_y0(operator, vector, options)
        if preconditioner is not None and not is_positive_semidefinite(preconditioner):
            raise ValueError("The preconditioner must be positive definite.")
        leaves = tr.tree_leaves(vector)
        size = sum(_numel(l) for l in leaves)
        max_steps = 10 * size if self.max_steps is None else self.max_steps  # cg.py:124-127
        flags = (nat.NSD if is_nsd else 0) | (0 if self.max_steps is None else nat.MAXSTEPS_GIVEN)
        se = 0 if self.stabilise_every is None else int(self.stabilise_every)
        b = ravel_leaves(leaves)
        y0f = None if y0 is None else ravel_leaves(tr.tree_leaves(y0))
        m = None if preconditioner is None else preconditioner.as_matrix()
        x, result, steps = _ops.cg(operator.as_matrix(), b, m, y0f, float(self.rtol), float(self.atol),
                                   int(max_steps), se, flags)
        solution = unravel_like(x, tr.struct_of(vector))
        return solution, result, {"num_steps": steps, "max_steps": self.max_steps}

    def transpose(self, state, options):
        transpose_options = {}
        if "preconditioner" in options:
            transpose_options["preconditioner"] = options["preconditioner"].transpose()
        psd_op, is_nsd = state
        return (psd_op.transpose(), is_nsd), transpose_options

    def conj(self, state, options):
        conj_options = {}
        if "preconditioner" in options:
            conj_options["preconditioner"] = conj(options["preconditioner"])
        psd_op, is_nsd = state
        return (conj(psd_op), is_nsd), conj_options

    def assume_full_rank(self):
        return True


def _numel(t):
    import math

    return math.prod(t.shape)


def NormalCG(*args, **kwargs):
    """Deprecated helper (lineax/_solver/cg.py:271-284). Use `Normal(CG(...))`."""
    from .normal import Normal

    warnings.warn("`NormalCG(...)` is deprecated in favour of `Normal(CG(...))`.",
                  DeprecationWarning, stacklevel=2)
    return Normal(CG(*args, **kwargs))
