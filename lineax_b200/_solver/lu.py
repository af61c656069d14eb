"""`LU` solver, drop-in for lineax/_solver/lu.py:37-98 on the native kernels (csrc/lu.cu)."""
from __future__ import annotations

import torch

from .. import _ops
from .._operator import diagonal, is_diagonal
from .._solution import RESULTS
from .._solve import AbstractLinearSolver
from .misc import pack_structures, ravel_vector, transpose_packed_structures, unravel_solution


class LU(AbstractLinearSolver):
    """LU solver for linear systems (square, nonsingular operators).

    state = ((lu, piv), packed_structures, transposed)  -- lineax/_solver/lu.py:34,54.
    `piv` is the 0-based int32 row-swap sequence of LAPACK getrf.
    """

    def init(self, operator, options):
        del options
        if operator.in_size() != operator.out_size():
            raise ValueError("`LU` may only be used for linear solves with square matrices")
        packed_structures = pack_structures(operator)
        if is_diagonal(operator):  # lu.py:50-51
            mat = operator.as_matrix()
            lu = mat, torch.arange(operator.in_size(), dtype=torch.int32, device=mat.device)
        else:
            lu = tuple(_ops.lu_factor(operator.as_matrix()))
        return lu, packed_structures, False

    def compute(self, state, vector, options):
        del options
        (lu, piv), packed_structures, transpose = state
        vector = ravel_vector(vector, packed_structures)
        solution = _ops.lu_solve(lu, piv, vector, bool(transpose))
        solution = unravel_solution(solution, packed_structures)
        return solution, RESULTS.successful, {}

    def _fused(self, operator, vector, options, keep_state: bool):
        """init + compute in one kernel: A is read from HBM once and the factors never
        round-trip through HBM unless the caller wants the state."""
        if operator.in_size() != operator.out_size():
            raise ValueError("`LU` may only be used for linear solves with square matrices")
        if is_diagonal(operator):
            return None
        packed_structures = pack_structures(operator)
        flat = ravel_vector(vector, packed_structures)
        x, lu, piv = _ops.lu_factor_solve(operator.as_matrix(), flat, keep_state)
        state = ((lu, piv), packed_structures, False) if keep_state else None
        return unravel_solution(x, packed_structures), RESULTS.successful, {}, state

    def transpose(self, state, options):
        lu_and_piv, packed_structures, transpose = state
        return (lu_and_piv, transpose_packed_structures(packed_structures), not transpose), {}

    def conj(self, state, options):
        (lu, piv), packed_structures, transpose = state
        return ((lu.conj(), piv), packed_structures, not transpose), {}

    def assume_full_rank(self):
        return True
