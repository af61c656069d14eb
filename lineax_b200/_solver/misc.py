"""Solver utilities mirroring lineax/_solver/misc.py:30-114 (host side; defines the flat
layout the kernels see: leaves concatenated in tree_leaves order, promoted to result_type)."""
from __future__ import annotations

import math
from typing import Any

import torch

from .. import _tree as tr
from .._operator import AbstractLinearOperator, IdentityLinearOperator, linearise


def preconditioner_and_y0(operator, vector, options: dict):
    """lineax/_solver/misc.py:30-60. Returns (preconditioner or None for identity, y0 or None for zeros)."""
    structure = operator.in_structure()
    preconditioner = options.get("preconditioner", None)
    if preconditioner is not None:
        preconditioner = linearise(preconditioner)
        if not isinstance(preconditioner, AbstractLinearOperator):
            raise ValueError("The preconditioner must be a linear operator.")
        if not tr.structure_equal(preconditioner.in_structure(), structure):
            raise ValueError(
                "The preconditioner must have `in_structure` that matches the "
                "operator's `in_strucure`."
            )
        if not tr.structure_equal(preconditioner.out_structure(), structure):
            raise ValueError(
                "The preconditioner must have `out_structure` that matches the "
                "operator's `in_structure`."
            )
        if isinstance(preconditioner, IdentityLinearOperator):
            preconditioner = None
    y0 = options.get("y0", None)
    if y0 is not None:
        y0 = tr.tree_map(tr.inexact_asarray, y0)
        if not tr.structure_equal(tr.struct_of(y0), tr.struct_of(vector)):
            raise ValueError("`y0` must have the same structure, shape, and dtype as `vector`")
    return preconditioner, y0


class PackedStructures:
    """(out_structure, in_structure) of the operator (lineax/_solver/misc.py:70-76)."""

    def __init__(self, out_structure, in_structure):
        self.out_structure = out_structure
        self.in_structure = in_structure


def pack_structures(operator) -> PackedStructures:
    return PackedStructures(operator.out_structure(), operator.in_structure())


def transpose_packed_structures(ps: PackedStructures) -> PackedStructures:
    return PackedStructures(ps.in_structure, ps.out_structure)


def ravel_leaves(leaves) -> torch.Tensor:
    dtype = tr.result_type(*leaves)
    return torch.cat([x.to(dtype).reshape(-1) for x in leaves])


def ravel_vector(pytree, packed_structures: PackedStructures) -> torch.Tensor:
    """lineax/_solver/misc.py:79-90."""
    if not tr.structure_equal(tr.struct_of(pytree), packed_structures.out_structure):
        raise ValueError("pytree does not match out_structure")
    return ravel_leaves(tr.tree_leaves(pytree))


def unravel_like(flat: torch.Tensor, structure) -> Any:
    leaves, treedef = tr.tree_flatten(structure)
    sizes = [math.prod(s.shape) for s in leaves]
    parts = torch.split(flat, sizes, dim=-1) if len(sizes) else ()
    shaped = [p.reshape(s.shape).to(s.dtype) for p, s in zip(parts, leaves)]
    return tr.tree_unflatten(shaped, treedef)


def unravel_solution(solution: torch.Tensor, packed_structures: PackedStructures):
    """lineax/_solver/misc.py:93-105: split, reshape and cast back per leaf."""
    return unravel_like(solution, packed_structures.in_structure)
