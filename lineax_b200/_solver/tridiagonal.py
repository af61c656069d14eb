"""`Tridiagonal`, drop-in for lineax/_solver/tridiagonal.py:36-88 on csrc/tridiagonal.cu."""
from __future__ import annotations

from .. import _ops
from .._operator import is_tridiagonal, tridiagonal
from .._solution import RESULTS
from .._solve import AbstractLinearSolver
from .misc import pack_structures, ravel_vector, transpose_packed_structures, unravel_solution


class Tridiagonal(AbstractLinearSolver):
    """Tridiagonal solver: Gaussian elimination with partial pivoting (LAPACK gtsv semantics).

    state = ((diagonal, lower_diagonal, upper_diagonal), packed_structures)  -- tridiagonal.py:33,52.
    """

    def init(self, operator, options):
        del options
        if operator.in_size() != operator.out_size():
            raise ValueError("`Tridiagonal` may only be used for linear solves with square matrices")
        if not is_tridiagonal(operator):
            raise ValueError("`Tridiagonal` may only be used for linear solves with tridiagonal matrices")
        return tridiagonal(operator), pack_structures(operator)

    def compute(self, state, vector, options):
        (diagonal, lower_diagonal, upper_diagonal), packed_structures = state
        del options
        vector = ravel_vector(vector, packed_structures)
        solution = _ops.tridiagonal_solve(diagonal, lower_diagonal, upper_diagonal, vector)
        return unravel_solution(solution, packed_structures), RESULTS.successful, {}

    def transpose(self, state, options):
        (diagonal, lower_diagonal, upper_diagonal), packed_structures = state
        return ((diagonal, upper_diagonal, lower_diagonal),
                transpose_packed_structures(packed_structures)), options

    def conj(self, state, options):
        (d, l, u), packed_structures = state
        return ((d.conj(), l.conj(), u.conj()), packed_structures), options

    def assume_full_rank(self):
        return True
