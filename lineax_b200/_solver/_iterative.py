"""Shared host-side plumbing of the iterative solvers (option parsing, flags, flat layout)."""
from __future__ import annotations

import math

from .. import _native as nat
from .. import _tree as tr
from .misc import preconditioner_and_y0, ravel_leaves, unravel_like


def check_tols(self):
    """`__check_init__` of the reference's iterative solvers (e.g. cg.py:75-86)."""
    if isinstance(self.rtol, (int, float)) and self.rtol < 0:
        raise ValueError("Tolerances must be non-negative.")
    if isinstance(self.atol, (int, float)) and self.atol < 0:
        raise ValueError("Tolerances must be non-negative.")
    if isinstance(self.atol, (int, float)) and isinstance(self.rtol, (int, float)):
        if self.atol == 0 and self.rtol == 0 and self.max_steps is None:
            raise ValueError(
                "Must specify `rtol`, `atol`, or `max_steps` (or some combination of all three)."
            )


def flat_problem(operator, vector, options):
    """-> (A, b, M or None, y0 or None, size)."""
    preconditioner, y0 = preconditioner_and_y0(operator, vector, options)
    leaves = tr.tree_leaves(vector)
    size = sum(math.prod(l.shape) for l in leaves)
    b = ravel_leaves(leaves)
    y0f = None if y0 is None else ravel_leaves(tr.tree_leaves(y0))
    m = None if preconditioner is None else preconditioner.as_matrix()
    return operator.as_matrix(), b, m, y0f, size, preconditioner


def steps_flags(max_steps, size):
    ms = 10 * size if max_steps is None else int(max_steps)  # "Copied from SciPy!", cg.py:124-127
    return ms, (0 if max_steps is None else nat.MAXSTEPS_GIVEN)


def transpose_options(options):
    out = {}
    if "preconditioner" in options:
        out["preconditioner"] = options["preconditioner"].transpose()
    return out


def conj_options(options):
    from .._operator import conj

    out = {}
    if "preconditioner" in options:
        out["preconditioner"] = conj(options["preconditioner"])
    return out
