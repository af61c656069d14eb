"""PyTree plumbing (torch.utils._pytree) + structure helpers used by operators/solvers."""
from __future__ import annotations

import math
from typing import Any

import torch
import torch.utils._pytree as pytree

tree_flatten = pytree.tree_flatten
tree_unflatten = pytree.tree_unflatten
tree_map = pytree.tree_map
tree_leaves = pytree.tree_leaves
tree_structure = pytree.tree_structure


class ShapeDtypeStruct:
    """Shape/dtype of one leaf (role of jax.ShapeDtypeStruct)."""

    __slots__ = ("shape", "dtype")

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype

    @property
    def size(self):
        return math.prod(self.shape)

    @property
    def ndim(self):
        return len(self.shape)

    def __eq__(self, other):
        return (isinstance(other, ShapeDtypeStruct) and self.shape == other.shape
                and self.dtype == other.dtype)

    def __hash__(self):
        return hash((self.shape, self.dtype))

    def __repr__(self):
        return f"ShapeDtypeStruct(shape={self.shape}, dtype={self.dtype})"


_default_device = None


def default_device() -> torch.device:
    global _default_device
    if _default_device is None:
        _default_device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    return _default_device


def set_default_device(device):
    global _default_device
    _default_device = torch.device(device)


def default_floating_dtype():
    return torch.get_default_dtype()


def inexact_asarray(x, device=None) -> torch.Tensor:
    """lineax/_misc.py:82-86: arrays of inexact dtype; other leaves promoted to the default float."""
    if isinstance(x, torch.Tensor):
        t = x
        if not t.is_cuda and device is None and default_device().type == "cuda":
            t = t.to(default_device())  # host arrays are accepted and staged to the GPU
        elif device is not None and t.device != torch.device(device):
            t = t.to(device)
    else:
        import numpy as np

        if isinstance(x, np.ndarray) or isinstance(x, np.generic):
            t = torch.as_tensor(np.asarray(x))
        else:
            t = torch.as_tensor(x)
        t = t.to(device if device is not None else default_device())
    if not (t.is_floating_point() or t.is_complex()):
        t = t.to(default_floating_dtype())
    return t


def struct_of(tree) -> Any:
    """eval_shape of a PyTree of tensors."""
    return tree_map(lambda t: ShapeDtypeStruct(t.shape, t.dtype), tree)


def structure_equal(a, b) -> bool:
    la, ta = tree_flatten(a)
    lb, tb = tree_flatten(b)
    return ta == tb and la == lb


def tree_size(struct) -> int:
    return sum(s.size if isinstance(s, ShapeDtypeStruct) else s.numel() for s in tree_leaves(struct))


def result_type(*leaves):
    dts = [(l.dtype) for l in leaves]
    out = dts[0]
    for d in dts[1:]:
        out = torch.promote_types(out, d)
    return out
