"""Norms mirroring lineax/_norm.py.  Inside the solvers these are fused into the kernels;
the standalone functions identify which test a solver was asked for (only the defaults are
native: `max_norm` for CG/BiCGStab/GMRES, `two_norm` for LSMR) and are callable on tensors."""
from __future__ import annotations

import torch

from . import _tree as tr


def two_norm(x):
    """sqrt(sum x_i^2) over a PyTree (lineax/_norm.py:59-82)."""
    from . import _ops

    leaves = [tr.inexact_asarray(l) for l in tr.tree_leaves(x)]
    if sum(l.numel() for l in leaves) == 0:
        return torch.zeros((), dtype=tr.default_floating_dtype(), device=tr.default_device())
    flat = torch.cat([l.reshape(-1) for l in leaves])
    return _ops.norms(flat)[0]


def max_norm(x):
    """max |x_i| over a PyTree, NaN-propagating (lineax/_norm.py:123-139)."""
    from . import _ops

    leaves = [tr.inexact_asarray(l) for l in tr.tree_leaves(x)]
    if sum(l.numel() for l in leaves) == 0:
        return torch.zeros((), dtype=tr.default_floating_dtype(), device=tr.default_device())
    flat = torch.cat([l.reshape(-1) for l in leaves])
    return _ops.norms(flat)[1]


def rms_norm(x):
    """two_norm / sqrt(size) (lineax/_norm.py:104-120)."""
    import math

    size = sum(tr.inexact_asarray(l).numel() for l in tr.tree_leaves(x))
    if size == 0:
        return torch.zeros((), dtype=tr.default_floating_dtype(), device=tr.default_device())
    return two_norm(x) / math.sqrt(size)


def tree_dot(a, b):
    """sum conj(a) b over matching PyTrees (lineax/_norm.py:27-47)."""
    from . import _ops

    la = [tr.inexact_asarray(l).reshape(-1) for l in tr.tree_leaves(a)]
    lb = [tr.inexact_asarray(l).reshape(-1) for l in tr.tree_leaves(b)]
    if tr.tree_structure(a) != tr.tree_structure(b):
        raise ValueError("trees must have the same structure")
    if len(la) == 0:
        return torch.zeros((), dtype=tr.default_floating_dtype(), device=tr.default_device())
    return _ops.dot(torch.cat(la), torch.cat(lb))
