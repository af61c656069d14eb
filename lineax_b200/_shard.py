"""Multi-GPU batch sharding: independent vmapped systems are split into
contiguous blocks, one block per rank (one process per GPU), with NO data-path collective.
The only communication is the optional gather of results at the end."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int) -> list:
    """Contiguous, balanced block boundaries: ranks [0, batch % world) get one extra system."""
    base, extra = divmod(batch, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return bounds


def local_slice(batch: int, rank: Optional[int] = None, world: Optional[int] = None) -> slice:
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    b = shard_bounds(batch, world)
    return slice(b[rank], b[rank + 1])


def gather_batch(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """all_gather the per-rank blocks of a batch-sharded result back into [batch, ...]."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    bounds = shard_bounds(batch, world)
    width = max(bounds[r + 1] - bounds[r] for r in range(world))
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][: bounds[r + 1] - bounds[r]] for r in range(world)], dim=0)


def solve_sharded(fn, *batched_args, gather: bool = True, group=None):
    """Run `fn(*args)` (a vmapped solve) on this rank's contiguous block of the leading batch
    dimension and optionally gather the per-rank results. Every rank passes the same full-batch
    arguments (or arguments already resident on its device)."""
    batch = batched_args[0].shape[0]
    sl = local_slice(batch, None if group is None else dist.get_rank(group),
                     None if group is None else dist.get_world_size(group))
    out = fn(*[a[sl] for a in batched_args])
    if not gather:
        return out
    if isinstance(out, tuple):
        return tuple(gather_batch(o, batch, group) for o in out)
    return gather_batch(out, batch, group)
