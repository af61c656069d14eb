"""`QR`, drop-in for lineax/_solver/qr.py:37-118 on csrc/direct.cu (geqrf / ormqr / trtrs)."""
from __future__ import annotations

from .. import _ops
from .._solution import RESULTS
from .._solve import AbstractLinearSolver
from .misc import pack_structures, ravel_vector, transpose_packed_structures, unravel_solution


class QR(AbstractLinearSolver):
    """QR solver; handles non-square, full-rank operators (least squares / minimum norm).

    state = ((a, taus), transpose, packed_structures)  -- qr.py:34,65; `a` is geqrf's output
    of A (or of A^T when wide): R in the upper triangle, Householder vectors below.
    """

    def init(self, operator, options):
        del options
        from .gmres import _is_row_sharded

        if _is_row_sharded(operator):  # TSQR over the row blocks (lineax_b200.distributed.RowShardedQR)
            return operator, False, None
        matrix = operator.as_matrix()
        m, n = matrix.shape
        transpose = n > m
        a, taus = _ops.qr_factor(matrix)  # factors A^T itself when wide (qr.py:59-61)
        return (a, taus), transpose, pack_structures(operator)

    def compute(self, state, vector, options):
        from .gmres import _is_row_sharded

        if _is_row_sharded(state[0]):
            from ..distributed import RowShardedQR

            op = state[0]
            solver = op.sharded_solver("qr", (op.rows, op.cols, op.local.dtype), lambda: RowShardedQR(
                op.rows, op.cols, dtype=op.local.dtype, group=op.group))
            return solver.solve(op.local, vector), RESULTS.successful, {}
        (a, taus), transpose, packed_structures = state
        del options
        vector = ravel_vector(vector, packed_structures)
        solution = _ops.qr_solve(a, taus, vector, bool(transpose))
        return unravel_solution(solution, packed_structures), RESULTS.successful, {}

    def transpose(self, state, options):
        (a, taus), transpose, structures = state
        return ((a, taus), not transpose, transpose_packed_structures(structures)), {}

    def conj(self, state, options):
        (a, taus), transpose, structures = state
        return ((a.conj(), taus.conj()), transpose, structures), {}

    def assume_full_rank(self):
        return True
