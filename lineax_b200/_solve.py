"""`linear_solve`, the solver ABC and `AutoLinearSolver`, mirroring lineax/_solve.py.

Dispatch is the reference's: validate structures -> `solver.init` -> `solver.compute` ->
non-finite result rewriting (lineax/_solve.py:97-129) -> `Solution`.  The impl rule of the
reference's primitive is literally `solver.compute(state, vector, options)` (_solve.py:97-98);
here that call lands in a native kernel.  Batching (`jax.vmap`) is `torch.func.vmap`: every
native op declares its batching rule (lineax_b200/_ops.py).
"""
from __future__ import annotations

import abc
from typing import Any, Optional

import torch

from . import _tree as tr
from ._operator import (
    AbstractLinearOperator, IdentityLinearOperator, is_diagonal, is_lower_triangular,
    is_negative_semidefinite, is_positive_semidefinite, is_tridiagonal, is_upper_triangular,
)
from ._solution import RESULTS, LinearSolveError, Solution

sentinel = object()


class AbstractLinearSolver(abc.ABC):
    """Abstract base class for all linear solvers (lineax/_solve.py:343-480)."""

    @abc.abstractmethod
    def init(self, operator: AbstractLinearOperator, options: dict) -> Any:
        """Do any initial computation on just the `operator` (e.g. factorise it)."""

    @abc.abstractmethod
    def compute(self, state: Any, vector, options: dict):
        """Solve against `vector`; returns `(solution, RESULTS code(s), stats dict)`."""

    @abc.abstractmethod
    def transpose(self, state: Any, options: dict):
        """State/options of the transposed operator; must equal `init(operator.T)`."""

    @abc.abstractmethod
    def conj(self, state: Any, options: dict):
        """State/options of the conjugated operator."""

    @abc.abstractmethod
    def assume_full_rank(self) -> bool:
        """Whether the solver assumes a full-rank operator (skips pseudo-inverse JVP terms)."""

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __hash__(self):
        return hash((type(self), tuple(sorted((k, repr(v)) for k, v in self.__dict__.items()))))

    def __repr__(self):
        args = ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items())
        return f"{type(self).__name__}({args})"


_qr_token, _diagonal_token, _well_posed_diagonal_token = "qr_token", "diagonal_token", "well_posed_diagonal_token"
_tridiagonal_token, _triangular_token, _cholesky_token = "tridiagonal_token", "triangular_token", "cholesky_token"
_lu_token, _svd_token = "lu_token", "svd_token"


def _lookup(token) -> AbstractLinearSolver:
    from . import _solver

    if token == _svd_token:
        raise NotImplementedError(
            "AutoLinearSolver(well_posed=False) dispatches to SVD for non-diagonal operators; "
            "SVD is outside the accelerated hot path (SURVEY.md section 2.1 #8). Use QR() or LSMR()."
        )
    return {
        _qr_token: lambda: _solver.QR(),
        _diagonal_token: lambda: _solver.Diagonal(),
        _well_posed_diagonal_token: lambda: _solver.Diagonal(well_posed=True),
        _tridiagonal_token: lambda: _solver.Tridiagonal(),
        _triangular_token: lambda: _solver.Triangular(),
        _cholesky_token: lambda: _solver.Cholesky(),
        _lu_token: lambda: _solver.LU(),
    }[token]()


class AutoLinearSolver(AbstractLinearSolver):
    """Chooses a solver from the operator's structure (lineax/_solve.py:518-645)."""

    def __init__(self, well_posed: Optional[bool]):
        self.well_posed = well_posed

    def _select_solver(self, operator):
        if self.well_posed is True:
            if operator.in_size() != operator.out_size():
                raise ValueError(
                    "Cannot use `AutoLinearSolver(well_posed=True)` with a non-square "
                    "operator. If you are trying solve a least-squares problem then "
                    "you should pass `solver=AutoLinearSolver(well_posed=False)`. By "
                    "default `linear_solve` assumes that the operator is "
                    "square and nonsingular."
                )
            if is_diagonal(operator):
                return _well_posed_diagonal_token
        elif self.well_posed is False:
            return _diagonal_token if is_diagonal(operator) else _svd_token
        elif self.well_posed is None:
            if operator.in_size() != operator.out_size():
                return _qr_token
            if is_diagonal(operator):
                return _diagonal_token
        else:
            raise ValueError(f"Invalid value `well_posed={self.well_posed}`.")
        if is_tridiagonal(operator):
            return _tridiagonal_token
        if is_lower_triangular(operator) or is_upper_triangular(operator):
            return _triangular_token
        if is_positive_semidefinite(operator) or is_negative_semidefinite(operator):
            return _cholesky_token
        return _lu_token

    def select_solver(self, operator) -> AbstractLinearSolver:
        return _lookup(self._select_solver(operator))

    def init(self, operator, options):
        token = self._select_solver(operator)
        return token, _lookup(token).init(operator, options)

    def compute(self, state, vector, options):
        token, state = state
        solution, result, _ = _lookup(token).compute(state, vector, options)
        return solution, result, {}

    def _fused(self, operator, vector, options, keep_state):
        token = self._select_solver(operator)
        inner = _lookup(token)
        if not hasattr(inner, "_fused"):
            return None
        out = inner._fused(operator, vector, options, keep_state)
        if out is None:
            return None
        solution, result, _, state = out
        return solution, result, {}, (None if state is None else (token, state))

    def transpose(self, state, options):
        token, state = state
        t_state, t_options = _lookup(token).transpose(state, options)
        return (token, t_state), t_options

    def conj(self, state, options):
        token, state = state
        c_state, c_options = _lookup(token).conj(state, options)
        return (token, c_state), c_options

    def assume_full_rank(self):
        return self.well_posed is not False


class _Config:
    """Global switches.
    keep_state: write solver state (e.g. LU factors) even when `linear_solve` fused init+compute;
                False lets `Solution.state` be recomputed lazily on first access instead.
    enable_x64: mirrors `jax_enable_x64` for BiCGStab's breakdown test (bicgstab.py:110);
                None = infer from the operand dtype."""

    keep_state = False
    enable_x64: Optional[bool] = None


config = _Config()


def _in_vmap(*trees) -> bool:
    try:
        from torch._C._functorch import is_batchedtensor
    except Exception:  # pragma: no cover
        return False
    return any(isinstance(l, torch.Tensor) and is_batchedtensor(l) for t in trees for l in tr.tree_leaves(t))


def _finalise(solution, result, vector, throw: bool):
    """lineax/_solve.py:104-128: rewrite results for non-finite output/input, then throw."""
    from . import _ops
    from ._solver.misc import ravel_leaves

    sol_leaves = tr.tree_leaves(solution)
    vec_leaves = tr.tree_leaves(vector)
    dev = (sol_leaves or vec_leaves)[0].device
    if not isinstance(result, torch.Tensor):
        result = torch.full((), int(result), dtype=torch.int32, device=dev)
    result = _ops.postprocess(ravel_leaves(sol_leaves), ravel_leaves(vec_leaves), result)
    if throw:
        _ops.throw_if_failed(result)
    return result


def linear_solve(operator, vector, solver: AbstractLinearSolver = AutoLinearSolver(well_posed=True),
                 *, options: Optional[dict] = None, state: Any = sentinel, throw: bool = True) -> Solution:
    """Solves a linear system `operator @ x = vector` (lineax/_solve.py:656-806).

    Same arguments and failure behaviour as `lineax.linear_solve`: `throw=True` raises
    (`LinearSolveError`) when `result != RESULTS.successful`; `throw=False` reports through
    `Solution.result`.  `state=` reuses a previous `solver.init(operator, options)`.
    """
    if isinstance(operator, torch.Tensor):
        raise ValueError(
            "`linear_solve(operator=...)` should be an `AbstractLinearOperator`, not a raw "
            "array. If you are trying to pass a matrix then this should be passed as "
            "`MatrixLinearOperator(matrix)`."
        )
    if options is None:
        options = {}
    vector = tr.tree_map(tr.inexact_asarray, vector)
    vector_struct = tr.struct_of(vector)
    if not tr.structure_equal(vector_struct, operator.out_structure()):
        raise ValueError(
            "Vector and operator structures do not match. Got a vector with structure "
            f"{vector_struct} and an operator with out-structure {operator.out_structure()}"
        )
    if isinstance(operator, IdentityLinearOperator):  # _solve.py:778-784
        dev = tr.tree_leaves(vector)[0].device
        return Solution(value=vector, result=torch.zeros((), dtype=torch.int32, device=dev),
                        stats={}, state=None if state is sentinel else state)
    fused = None
    if state is sentinel and hasattr(solver, "_fused"):
        fused = solver._fused(operator, vector, options, config.keep_state)
    if fused is not None:
        solution, result, stats, st = fused
        thunk = None if st is not None else (lambda: solver.init(operator, options))
        result = _finalise(solution, result, vector, throw)
        return Solution(value=solution, result=result, stats=stats, state=st, state_thunk=thunk)
    if state is sentinel:
        state = solver.init(operator, options)  # _solve.py:785-790
    solution, result, stats = solver.compute(state, vector, options)  # _solve.py:98
    result = _finalise(solution, result, vector, throw)
    return Solution(value=solution, result=result, stats=stats, state=state)


def invert(operator, solver: AbstractLinearSolver = AutoLinearSolver(well_posed=True)):
    """Operator whose `mv` solves against `operator` (lineax/_solve.py:809-871), reusing one `init`."""
    state = solver.init(operator, {})

    class _Inverse(AbstractLinearOperator):
        def mv(self, vector):
            return linear_solve(operator, vector, solver, state=state).value

        def as_matrix(self):
            raise NotImplementedError("materialising an inverse operator is not supported")

        def transpose(self):
            return invert(operator.transpose(), solver)

        def in_structure(self):
            return operator.out_structure()

        def out_structure(self):
            return operator.in_structure()

    return _Inverse()
