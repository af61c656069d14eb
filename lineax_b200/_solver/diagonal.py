"""`Diagonal`, drop-in for lineax/_solver/diagonal.py:36-105 (AutoLinearSolver dispatch target)."""
from __future__ import annotations

import torch

from .. import _ops
from .._operator import diagonal, has_unit_diagonal, is_diagonal
from .._solution import RESULTS
from .._solve import AbstractLinearSolver
from .misc import pack_structures, ravel_vector, transpose_packed_structures, unravel_solution


class Diagonal(AbstractLinearSolver):
    """Diagonal solver: elementwise division; with `well_posed=False` entries below
    `rcond * max|d|` are treated as zero (pseudo-inverse).  state = (diag or None, packed_structures)."""

    def __init__(self, well_posed: bool = False, rcond=None):
        self.well_posed, self.rcond = well_posed, rcond

    def init(self, operator, options):
        del options
        if operator.in_size() != operator.out_size():
            raise ValueError("`Diagonal` may only be used for linear solves with square matrices")
        if not is_diagonal(operator):
            raise ValueError("`Diagonal` may only be used for linear solves with diagonal matrices")
        packed_structures = pack_structures(operator)
        if has_unit_diagonal(operator):
            return None, packed_structures
        return diagonal(operator), packed_structures

    def compute(self, state, vector, options):
        diag, packed_structures = state
        del options
        vector = ravel_vector(vector, packed_structures)
        if diag is None:
            solution = vector
        else:
            if self.well_posed:
                rcond = -1.0
            else:  # resolve_rcond, lineax/_misc.py:30-38
                eps = torch.finfo(diag.dtype).eps
                size = diag.shape[-1]
                rcond = 2 * eps * size if self.rcond is None else (eps if self.rcond < 0 else float(self.rcond))
            solution = _ops.diagonal_solve(diag, vector, float(rcond))
        return unravel_solution(solution, packed_structures), RESULTS.successful, {}

    def transpose(self, state, options):
        del options
        diag, packed_structures = state
        return (diag, transpose_packed_structures(packed_structures)), {}

    def conj(self, state, options):
        del options
        diag, packed_structures = state
        return (None if diag is None else diag.conj(), packed_structures), {}

    def assume_full_rank(self):
        return self.well_posed
