"""RESULTS codes and the Solution container, mirroring lineax/_solution.py:52-88."""
from __future__ import annotations

from typing import Any

import torch

_singular_msg = (
    "A linear solver returned non-finite (NaN or inf) output. This usually means that an\n"
    "operator was not well-posed, and that its solver does not support this.\n\n"
    "If you are trying solve a linear least-squares problem then you should pass\n"
    "`solver=AutoLinearSolver(well_posed=False)`. By default `lineax_b200.linear_solve`\n"
    "assumes that the operator is square and nonsingular.\n\n"
    "If you *were* expecting this solver to work with this operator, then it may be because:\n\n"
    "(a) the operator is singular, and your code has a bug; or\n\n"
    "(b) the operator was nearly singular (i.e. it had a high condition number), and the\n"
    "    solver suffered from numerical instability issues; or\n\n"
    "(c) the operator is declared to exhibit a certain property (e.g. positive definiteness)\n"
    "    that is does not actually satisfy."
)
_nonfinite_msg = (
    "A linear solver received non-finite (NaN or inf) input and cannot determine a \n"
    "solution.\n\n"
    "This means that you have a bug upstream of Lineax and should check the inputs to\n"
    "`linear_solve` for non-finite values."
)


class _ResultsMeta(type):
    def __getitem__(cls, item):
        """`RESULTS[result]` -> human-readable message (as in lineax)."""
        code = int(item.item()) if isinstance(item, torch.Tensor) else int(item)
        return cls._messages[code]

    def __iter__(cls):
        return iter(range(len(cls._messages)))

    def __len__(cls):
        return len(cls._messages)


class RESULTS(metaclass=_ResultsMeta):
    """Integer-coded results, in lineax's definition order (lineax/_solution.py:52-68).

    Results are carried as int32 tensors (one code per system) so that they batch.
    """

    successful = 0
    max_steps_reached = 1
    singular = 2
    breakdown = 3
    stagnation = 4
    conlim = 5
    nonfinite_input = 6

    _names = (
        "successful", "max_steps_reached", "singular", "breakdown", "stagnation", "conlim",
        "nonfinite_input",
    )
    _messages = (
        "",
        "The maximum number of solver steps was reached. Try increasing `max_steps`.",
        _singular_msg,
        "A form of iterative breakdown has occured in a linear solve. Try using a different "
        "solver for this problem or increase `restart` if using GMRES.",
        "A stagnation in an iterative linear solve has occurred. Try increasing "
        "`stagnation_iters` or `restart`.",
        "Condition number of A seems to be larger than `conlim`.",
        _nonfinite_msg,
    )

    @staticmethod
    def where(pred, a, b):
        return torch.where(pred, torch.as_tensor(a), torch.as_tensor(b))

    @classmethod
    def name(cls, code) -> str:
        return cls._names[int(code)]


class LinearSolveError(RuntimeError):
    """Raised by `linear_solve(..., throw=True)` on failure (role of EquinoxRuntimeError)."""


class Solution:
    """The solution to a linear solve (lineax/_solution.py:71-88).

    Attributes: `value`, `result` (int32 tensor of RESULTS codes), `stats`, `state`.
    `state` is materialised lazily when the fused init+compute kernel skipped it.
    """

    def __init__(self, value, result, stats, state=None, state_thunk=None):
        self.value = value
        self.result = result
        self.stats = stats
        self._state = state
        self._state_thunk = state_thunk

    @property
    def state(self) -> Any:
        if self._state is None and self._state_thunk is not None:
            self._state = self._state_thunk()
            self._state_thunk = None
        return self._state

    def __repr__(self):
        return f"Solution(value={self.value!r}, result={self.result!r}, stats={self.stats!r})"
