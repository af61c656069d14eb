"""`LSMR`, drop-in for lineax/_solver/lsmr.py:55-424 on csrc/lsmr.cu."""
from __future__ import annotations

import math

import torch

from .. import _ops
from .. import _tree as tr
from .._norm import two_norm
from .._operator import conj, linearise
from .._solve import AbstractLinearSolver
from ._iterative import check_tols, steps_flags
from .misc import ravel_leaves, unravel_like


class LSMR(AbstractLinearSolver):
    """LSMR for any (non-square, singular) operator; returns the pseudo-inverse solution.

    Option: `y0`.  state = operator.  stats: num_steps, istop, norm_r, norm_Ar, norm_A,
    cond_A, norm_x (lsmr.py:334-342).
    """

    def __init__(self, rtol, atol, norm=two_norm, max_steps=None, conlim=1e8):
        self.rtol, self.atol, self.norm, self.max_steps, self.conlim = rtol, atol, norm, max_steps, conlim
        check_tols(self)
        if isinstance(conlim, (int, float)) and conlim < 0:
            raise ValueError("Tolerances must be non-negative.")
        if norm is not two_norm:
            raise NotImplementedError("the native LSMR kernel implements the default `two_norm` tests")

    def init(self, operator, options):
        from .gmres import _is_row_sharded

        return operator if _is_row_sharded(operator) else linearise(operator)

    def _compute_row_sharded(self, operator, vector, options):
        """ONE tall system row-partitioned over the GPUs (csrc/lsmr_dist.cu): `vector` is this rank's slice
        of b, the solution and the statistics are replicated on every rank."""
        from ..distributed import RowShardedLSMR

        key = (operator.rows, operator.cols, float(self.rtol), float(self.atol), float(self.conlim), self.max_steps,
               operator.local.dtype)
        solver = operator.sharded_solver("lsmr", key, lambda: RowShardedLSMR(
            operator.rows, operator.cols, float(self.rtol), float(self.atol), conlim=float(self.conlim),
            max_steps=self.max_steps, dtype=operator.local.dtype, group=operator.group))
        x, result, steps, stats = solver.solve(operator.local, vector, options.get("y0"))
        return x, result, stats

    def compute(self, state, vector, options):
        operator = state
        from .gmres import _is_row_sharded

        if _is_row_sharded(operator):
            return self._compute_row_sharded(operator, vector, options)
        a = operator.as_matrix()
        m, n = operator.out_size(), operator.in_size()
        min_dim = min(m, n)
        flags = 0
        if self.max_steps is None:  # lsmr.py:121-129, with the integer-overflow guard
            imax = torch.iinfo(torch.int32 if a.dtype == torch.float32 else torch.int64).max
            ms = imax if min_dim > imax / 10 else min_dim * 10
        else:
            ms, flags = steps_flags(self.max_steps, 0)
        b = ravel_leaves(tr.tree_leaves(vector))
        y0 = options.get("y0", None)
        y0f = None if y0 is None else ravel_leaves(tr.tree_leaves(tr.tree_map(tr.inexact_asarray, y0)))
        x, result, steps, st = _ops.lsmr(a, b, y0f, float(self.rtol), float(self.atol), float(self.conlim),
                                         int(ms), flags)
        stats = {
            "num_steps": steps, "istop": st[..., 0].to(torch.int32), "norm_r": st[..., 1],
            "norm_Ar": st[..., 2], "norm_A": st[..., 3], "cond_A": st[..., 4], "norm_x": st[..., 5],
        }
        return unravel_like(x, operator.in_structure()), result, stats

    def transpose(self, state, options):
        del options
        return state.transpose(), {}

    def conj(self, state, options):
        del options
        return conj(state), {}

    def assume_full_rank(self):
        return False
