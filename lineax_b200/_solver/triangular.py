"""`Triangular`, drop-in for lineax/_solver/triangular.py:43-115 (AutoLinearSolver dispatch target)."""
from __future__ import annotations

from .. import _ops
from .._operator import has_unit_diagonal, is_lower_triangular, is_upper_triangular
from .._solution import RESULTS
from .._solve import AbstractLinearSolver
from .misc import pack_structures, ravel_vector, transpose_packed_structures, unravel_solution


class Triangular(AbstractLinearSolver):
    """Triangular solver. state = (matrix, lower, unit_diagonal, packed_structures, transposed)."""

    def init(self, operator, options):
        del options
        if operator.in_size() != operator.out_size():
            raise ValueError("`Triangular` may only be used for linear solves with square matrices")
        if not (is_lower_triangular(operator) or is_upper_triangular(operator)):
            raise ValueError("`Triangular` may only be used for linear solves with triangular matrices")
        return (operator.as_matrix(), is_lower_triangular(operator), has_unit_diagonal(operator),
                pack_structures(operator), False)

    def compute(self, state, vector, options):
        matrix, lower, unit_diagonal, packed_structures, transpose = state
        del options
        vector = ravel_vector(vector, packed_structures)
        solution = _ops.triangular_solve(matrix, vector, bool(lower), bool(unit_diagonal), bool(transpose))
        return unravel_solution(solution, packed_structures), RESULTS.successful, {}

    def transpose(self, state, options):
        del options
        matrix, lower, unit_diagonal, packed_structures, transpose = state
        return (matrix, lower, unit_diagonal, transpose_packed_structures(packed_structures), not transpose), {}

    def conj(self, state, options):
        del options
        matrix, lower, unit_diagonal, packed_structures, transpose = state
        return (matrix.conj(), lower, unit_diagonal, packed_structures, transpose), {}

    def assume_full_rank(self):
        return True
