"""Materialised linear operators, mirroring the part of lineax/_operator.py on the hot path:
`MatrixLinearOperator` (232-285), `PyTreeLinearOperator` (335-485), `DiagonalLinearOperator`
(488-520), `TridiagonalLinearOperator` (826-887), `TaggedLinearOperator` (890-942),
`IdentityLinearOperator` (736-823) and the structure queries `is_*` (1587-1929),
`diagonal` (1394), `tridiagonal` (1476), `linearise`/`materialise`/`conj`.

Matrix-free / lazy operators (Jacobian, Function, Composed, Add, ...) have no materialised
buffer to hand a kernel and are out of scope.
Arrays are torch tensors (CUDA); under `torch.func.vmap` they are batched tensors and all
shapes below are the per-system (logical) shapes, exactly like under `jax.vmap`.
"""
from __future__ import annotations

import abc
import math
from typing import Any, Iterable

import torch

from . import _tree as tr
from ._tags import (
    diagonal_tag, lower_triangular_tag, negative_semidefinite_tag, positive_semidefinite_tag,
    symmetric_tag, transpose_tags, tridiagonal_tag, unit_diagonal_tag, upper_triangular_tag,
)
from ._tree import ShapeDtypeStruct


def _frozenset(x) -> frozenset:
    try:
        return frozenset(x)
    except TypeError:
        return frozenset([x])


class AbstractLinearOperator(abc.ABC):
    """Abstract base class for all linear operators (lineax/_operator.py:69-229)."""

    @abc.abstractmethod
    def mv(self, vector): ...

    @abc.abstractmethod
    def as_matrix(self) -> torch.Tensor: ...

    @abc.abstractmethod
    def transpose(self) -> "AbstractLinearOperator": ...

    @abc.abstractmethod
    def in_structure(self): ...

    @abc.abstractmethod
    def out_structure(self): ...

    def in_size(self) -> int:
        return tr.tree_size(self.in_structure())

    def out_size(self) -> int:
        return tr.tree_size(self.out_structure())

    @property
    def T(self):
        return self.transpose()

    def __neg__(self):
        return _scaled(self, -1.0)

    def __mul__(self, other):
        return _scaled(self, other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return _scaled(self, 1.0 / other)


def _scaled(op, c):
    """Eager (materialised) scaling; lineax builds a lazy Mul/Neg operator instead.
    NSD <-> PSD tags swap under negation (lineax/_operator.py wrapper rules 1935-2295)."""
    if not isinstance(c, (int, float)):
        raise ValueError("Can only multiply AbstractLinearOperators by scalars.")
    if isinstance(op, DiagonalLinearOperator):
        return DiagonalLinearOperator(tr.tree_map(lambda d: d * c, op.diagonal))
    if isinstance(op, TridiagonalLinearOperator):
        return TridiagonalLinearOperator(op.diagonal * c, op.lower_diagonal * c, op.upper_diagonal * c)
    tags = set(_tags_of(op))
    if c < 0:
        psd, nsd = positive_semidefinite_tag in tags, negative_semidefinite_tag in tags
        tags -= {positive_semidefinite_tag, negative_semidefinite_tag}
        if psd:
            tags.add(negative_semidefinite_tag)
        if nsd:
            tags.add(positive_semidefinite_tag)
    tags.discard(unit_diagonal_tag)
    if isinstance(op, TaggedLinearOperator):
        return TaggedLinearOperator(_scaled(op.operator, c), frozenset(tags))
    if isinstance(op, MatrixLinearOperator):
        return MatrixLinearOperator(op.matrix * c, frozenset(tags))
    if isinstance(op, PyTreeLinearOperator):
        return PyTreeLinearOperator(tr.tree_map(lambda x: x * c, op.pytree), op.out_structure(),
                                    frozenset(tags))
    raise NotImplementedError(f"scaling of {type(op).__name__}")


def _tags_of(op) -> frozenset:
    return getattr(op, "tags", frozenset())


def _matvec(matrix: torch.Tensor, vector: torch.Tensor) -> torch.Tensor:
    from . import _ops

    return _ops.matvec(matrix, vector, False)


class MatrixLinearOperator(AbstractLinearOperator):
    """Wraps a 2-dimensional array into a linear operator (lineax/_operator.py:232-285)."""

    def __init__(self, matrix, tags=()):
        matrix = tr.inexact_asarray(matrix)
        if matrix.ndim != 2:
            raise ValueError("`MatrixLinearOperator(matrix=...)` should be 2-dimensional.")
        self.matrix = matrix
        self.tags = _frozenset(tags)

    def mv(self, vector):
        return _matvec(self.matrix, tr.inexact_asarray(vector, self.matrix.device))

    def as_matrix(self):
        return self.matrix

    def transpose(self):
        if is_symmetric(self):
            return self
        return MatrixLinearOperator(self.matrix.mT, transpose_tags(self.tags))

    def in_structure(self):
        return ShapeDtypeStruct((self.matrix.shape[1],), self.matrix.dtype)

    def out_structure(self):
        return ShapeDtypeStruct((self.matrix.shape[0],), self.matrix.dtype)


class PyTreeLinearOperator(AbstractLinearOperator):
    """A PyTree of arrays as a linear operator (lineax/_operator.py:335-485).

    `pytree` has structure tree(out) o tree(in) with leaf (i, j) of shape
    `(*y_shape_j, *x_shape_i)`.  The kernels see its dense flattening `as_matrix()`:
    row blocks = out leaves, column blocks = in leaves (lineax/_operator.py:438-455).
    """

    def __init__(self, pytree, output_structure, tags=()):
        self.output_structure_ = tr.tree_map(
            lambda s: ShapeDtypeStruct(s.shape, s.dtype if s.dtype.is_floating_point or s.dtype.is_complex
                                       else tr.default_floating_dtype()),
            output_structure,
        )
        self.pytree = tr.tree_map(tr.inexact_asarray, pytree)
        self.tags = _frozenset(tags)
        out_leaves, out_def = tr.tree_flatten(self.output_structure_)
        # split pytree at the depth of the out-structure
        subtrees = out_def.flatten_up_to(self.pytree) if hasattr(out_def, "flatten_up_to") else None
        if subtrees is None:
            raise ValueError("`pytree` and `output_structure` are not consistent")
        in_struct = None
        for struct, sub in zip(out_leaves, subtrees):
            def get(leaf, struct=struct):
                nd = len(struct.shape)
                if tuple(leaf.shape[:nd]) != struct.shape:
                    raise ValueError("`pytree` and `output_structure` are not consistent")
                return ShapeDtypeStruct(leaf.shape[nd:], leaf.dtype)

            s = tr.tree_map(get, sub)
            if in_struct is None:
                in_struct = s
            elif not tr.structure_equal(in_struct, s):
                raise ValueError("`pytree` does not have a consistent `input_structure`")
        self.input_structure_ = in_struct
        self._out_def = out_def
        self._subtrees = subtrees

    def as_matrix(self):
        dtype = tr.result_type(*tr.tree_leaves(self.pytree))
        rows = []
        for struct, sub in zip(tr.tree_leaves(self.output_structure_), self._subtrees):
            leaves = tr.tree_leaves(sub)
            rows.append(torch.cat(
                [l.to(dtype).reshape(struct.size, math.prod(l.shape[struct.ndim:])) for l in leaves],
                dim=1))
        return torch.cat(rows, dim=0)

    def mv(self, vector):
        from ._solver.misc import ravel_leaves, unravel_like

        flat = ravel_leaves(tr.tree_leaves(tr.tree_map(tr.inexact_asarray, vector)))
        return unravel_like(_matvec(self.as_matrix(), flat), self.out_structure())

    def transpose(self):
        if is_symmetric(self):
            return self
        out_leaves = tr.tree_leaves(self.output_structure_)
        in_leaves, in_def = tr.tree_flatten(self.input_structure_)
        # transposed[i][j] = moveaxis(pytree[j][i]) with out dims moved last
        cols = [[] for _ in in_leaves]
        for struct, sub in zip(out_leaves, self._subtrees):
            nd = struct.ndim
            for i, leaf in enumerate(tr.tree_leaves(sub)):
                cols[i].append(leaf.movedim(list(range(nd)), list(range(-nd, 0))) if nd else leaf)
        new = in_def.unflatten([self._out_def.unflatten(c) for c in cols])
        return PyTreeLinearOperator(new, self.in_structure(), transpose_tags(self.tags))

    def in_structure(self):
        return self.input_structure_

    def out_structure(self):
        return self.output_structure_


class DiagonalLinearOperator(AbstractLinearOperator):
    """Diagonal operator storing only the diagonal (lineax/_operator.py:488-520)."""

    def __init__(self, diagonal):
        self.diagonal = tr.tree_map(tr.inexact_asarray, diagonal)

    def mv(self, vector):
        from . import _ops

        return tr.tree_map(lambda d, v: _ops.diag_mv(d, tr.inexact_asarray(v, d.device)),
                           self.diagonal, vector)

    def as_matrix(self):
        return torch.diag_embed(diagonal(self))

    def transpose(self):
        return self

    def in_structure(self):
        return tr.struct_of(self.diagonal)

    def out_structure(self):
        return tr.struct_of(self.diagonal)


class IdentityLinearOperator(AbstractLinearOperator):
    """Identity (lineax/_operator.py:736-823); `linear_solve` short-circuits it (_solve.py:778-784)."""

    def __init__(self, input_structure, output_structure=None):
        self.input_structure_ = input_structure
        self.output_structure_ = input_structure if output_structure is None else output_structure

    def mv(self, vector):
        if not tr.structure_equal(self.input_structure_, self.output_structure_):
            raise NotImplementedError("non-square IdentityLinearOperator.mv")
        return vector

    def as_matrix(self):
        leaves = tr.tree_leaves(self.input_structure_)
        return torch.eye(self.out_size(), self.in_size(), dtype=tr.result_type(*leaves),
                         device=tr.default_device())

    def transpose(self):
        return IdentityLinearOperator(self.output_structure_, self.input_structure_)

    def in_structure(self):
        return self.input_structure_

    def out_structure(self):
        return self.output_structure_


class TridiagonalLinearOperator(AbstractLinearOperator):
    """Tridiagonal operator from its three diagonals (lineax/_operator.py:826-887)."""

    def __init__(self, diagonal, lower_diagonal, upper_diagonal):
        self.diagonal = tr.inexact_asarray(diagonal)
        self.lower_diagonal = tr.inexact_asarray(lower_diagonal)
        self.upper_diagonal = tr.inexact_asarray(upper_diagonal)
        (size,) = self.diagonal.shape
        if tuple(self.lower_diagonal.shape) != (size - 1,):
            raise ValueError("lower_diagonal and diagonal do not have consistent size")
        if tuple(self.upper_diagonal.shape) != (size - 1,):
            raise ValueError("upper_diagonal and diagonal do not have consistent size")

    def mv(self, vector):
        from . import _ops

        return _ops.tridiag_mv(self.diagonal, self.lower_diagonal, self.upper_diagonal,
                               tr.inexact_asarray(vector, self.diagonal.device))

    def as_matrix(self):
        return (torch.diag_embed(self.diagonal) + torch.diag_embed(self.lower_diagonal, offset=-1)
                + torch.diag_embed(self.upper_diagonal, offset=1))

    def transpose(self):
        return TridiagonalLinearOperator(self.diagonal, self.upper_diagonal, self.lower_diagonal)

    def in_structure(self):
        return ShapeDtypeStruct(self.diagonal.shape, self.diagonal.dtype)

    def out_structure(self):
        return ShapeDtypeStruct(self.diagonal.shape, self.diagonal.dtype)


class TaggedLinearOperator(AbstractLinearOperator):
    """Wraps an operator and declares tags for it (lineax/_operator.py:890-942)."""

    def __init__(self, operator: AbstractLinearOperator, tags):
        self.operator = operator
        self.tags = _frozenset(tags)

    def mv(self, vector):
        return self.operator.mv(vector)

    def as_matrix(self):
        return self.operator.as_matrix()

    def transpose(self):
        return TaggedLinearOperator(self.operator.transpose(), transpose_tags(self.tags))

    def in_structure(self):
        return self.operator.in_structure()

    def out_structure(self):
        return self.operator.out_structure()


# ------------------------------------------------------------------ queries ----
_DENSE = (MatrixLinearOperator, PyTreeLinearOperator)


def _has_real_dtype(op) -> bool:
    leaves = tr.tree_leaves((op.in_structure(), op.out_structure()))
    return not tr.result_type(*leaves).is_complex


def _query(name, dense_rule, special):
    def q(operator) -> bool:
        if isinstance(operator, TaggedLinearOperator):
            tag = _TAG_FOR.get(name)
            if tag is not None and tag in operator.tags:
                return True
            if name in ("is_symmetric", "is_tridiagonal") and diagonal_tag in operator.tags:
                return True
            if name == "is_symmetric" and (
                positive_semidefinite_tag in operator.tags or negative_semidefinite_tag in operator.tags
            ) and _has_real_dtype(operator):
                return True
            return q(operator.operator)
        if isinstance(operator, _DENSE):
            return dense_rule(operator)
        for cls, val in special.items():
            if isinstance(operator, cls):
                return val(operator) if callable(val) else val
        raise NotImplementedError(f"`{name}` has not been implemented for {type(operator).__name__}")

    q.__name__ = name
    q.__doc__ = f"Structure query `{name}` (lineax/_operator.py:1587-1929)."
    return q


_TAG_FOR = {
    "is_symmetric": symmetric_tag, "is_diagonal": diagonal_tag, "is_tridiagonal": tridiagonal_tag,
    "has_unit_diagonal": unit_diagonal_tag, "is_lower_triangular": lower_triangular_tag,
    "is_upper_triangular": upper_triangular_tag, "is_positive_semidefinite": positive_semidefinite_tag,
    "is_negative_semidefinite": negative_semidefinite_tag,
}
_square = lambda op: tr.structure_equal(op.in_structure(), op.out_structure())

is_symmetric = _query(
    "is_symmetric",
    lambda op: symmetric_tag in op.tags or diagonal_tag in op.tags or (
        (positive_semidefinite_tag in op.tags or negative_semidefinite_tag in op.tags)
        and _has_real_dtype(op)),
    {IdentityLinearOperator: _square, DiagonalLinearOperator: True, TridiagonalLinearOperator: False},
)
is_diagonal = _query(
    "is_diagonal",
    lambda op: diagonal_tag in op.tags or (op.in_size() == 1 and op.out_size() == 1),
    {IdentityLinearOperator: True, DiagonalLinearOperator: True,
     TridiagonalLinearOperator: lambda op: op.in_size() == 1},
)
is_tridiagonal = _query(
    "is_tridiagonal",
    lambda op: tridiagonal_tag in op.tags or diagonal_tag in op.tags,
    {IdentityLinearOperator: True, DiagonalLinearOperator: True, TridiagonalLinearOperator: True},
)
has_unit_diagonal = _query(
    "has_unit_diagonal", lambda op: unit_diagonal_tag in op.tags,
    {IdentityLinearOperator: True, DiagonalLinearOperator: False, TridiagonalLinearOperator: False},
)
is_lower_triangular = _query(
    "is_lower_triangular", lambda op: lower_triangular_tag in op.tags,
    {IdentityLinearOperator: True, DiagonalLinearOperator: True, TridiagonalLinearOperator: False},
)
is_upper_triangular = _query(
    "is_upper_triangular", lambda op: upper_triangular_tag in op.tags,
    {IdentityLinearOperator: True, DiagonalLinearOperator: True, TridiagonalLinearOperator: False},
)
is_positive_semidefinite = _query(
    "is_positive_semidefinite", lambda op: positive_semidefinite_tag in op.tags,
    {IdentityLinearOperator: _square, DiagonalLinearOperator: False, TridiagonalLinearOperator: False},
)
is_negative_semidefinite = _query(
    "is_negative_semidefinite", lambda op: negative_semidefinite_tag in op.tags,
    {IdentityLinearOperator: False, DiagonalLinearOperator: False, TridiagonalLinearOperator: False},
)


def diagonal(operator) -> torch.Tensor:
    """Extract the diagonal as a vector (lineax/_operator.py:1394-1462)."""
    if isinstance(operator, TaggedLinearOperator):
        return diagonal(operator.operator)
    if isinstance(operator, DiagonalLinearOperator):
        leaves = tr.tree_leaves(operator.diagonal)
        dtype = tr.result_type(*leaves)
        return torch.cat([l.to(dtype).reshape(-1) for l in leaves])
    if isinstance(operator, TridiagonalLinearOperator):
        return operator.diagonal
    if isinstance(operator, IdentityLinearOperator):
        return torch.ones(operator.in_size(), device=tr.default_device())
    return torch.diagonal(operator.as_matrix(), 0, -2, -1)


def tridiagonal(operator):
    """(diagonal, lower, upper) (lineax/_operator.py:1476-1580)."""
    if isinstance(operator, TaggedLinearOperator):
        return tridiagonal(operator.operator)
    if isinstance(operator, TridiagonalLinearOperator):
        return operator.diagonal, operator.lower_diagonal, operator.upper_diagonal
    if isinstance(operator, (DiagonalLinearOperator, IdentityLinearOperator)):
        d = diagonal(operator)
        z = torch.zeros(d.shape[0] - 1, dtype=d.dtype, device=d.device)
        return d, z, z
    m = operator.as_matrix()
    return (torch.diagonal(m, 0, -2, -1), torch.diagonal(m, -1, -2, -1), torch.diagonal(m, 1, -2, -1))


def linearise(operator):
    """Materialised operators are already linear (lineax/_operator.py:1206-1260)."""
    return operator


def materialise(operator):
    """Already materialised (lineax/_operator.py:1263-1391)."""
    return operator


def conj(operator):
    """Complex conjugate; the native path is real-only so this is the identity
    (lineax/_operator.py:2301-2408)."""
    leaves = tr.tree_leaves((operator.in_structure(), operator.out_structure()))
    if tr.result_type(*leaves).is_complex:
        raise NotImplementedError("complex operators are outside the accelerated hot path")
    return operator
