"""Row-sharded solves of ONE large system across the GPUs of an NVLink box.

One process per GPU (`torch.distributed`, NCCL for the rendezvous only).  The data-path exchanges
(GMRES: all-gather of the Krylov vector, all-reduce of the Gram-Schmidt scalars; LSMR: all-reduce of
the partial `A^T u` and of `||u||^2`) are NOT NCCL calls: they are fused into the persistent solver
kernels over peer memory obtained from `torch.distributed._symmetric_memory`
(csrc/gmres_dist.cu, csrc/lsmr_dist.cu).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _native as nat
from . import _ops
from ._operator import AbstractLinearOperator
from ._shard import shard_bounds
from ._tree import ShapeDtypeStruct


class RowShardedGMRES:
    """Restarted GMRES (lineax/_solver/gmres.py semantics) on a row-partitioned dense operator.

    Every rank constructs it with the same arguments and then calls `solve(A_local, b_local)` with
    its contiguous block of rows (`row_range(rank)`); returns `(x_local, result, num_steps)`.
    """

    def __init__(self, n: int, rtol: float, atol: float, *, restart: int = 20, stagnation_iters: int = 20,
                 max_steps=None, dtype=torch.float32, group=None):
        import torch.distributed._symmetric_memory as symm

        self.n, self.rtol, self.atol = int(n), float(rtol), float(atol)
        self.restart, self.stagnation_iters, self.max_steps = int(restart), int(stagnation_iters), max_steps
        self.dtype, self.group = dtype, group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.bounds = shard_bounds(self.n, self.world)
        self.sfx = nat.suffix(dtype)
        self.device = torch.device("cuda", torch.cuda.current_device())
        nbytes = nat.fn(f"lxb_gmres_rowsharded_symm_bytes_{self.sfx}")(self.n)
        self.symm_buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.symm_buf.zero_()
        self.handle = symm.rendezvous(self.symm_buf, self.group)
        self.peers_dev = int(self.handle.buffer_ptrs_dev)
        torch.cuda.synchronize()
        dist.barrier(self.group)  # every rank's flags are zero before the first kernel starts

    def row_range(self, rank=None):
        r = self.rank if rank is None else rank
        return self.bounds[r], self.bounds[r + 1]

    def solve(self, a_local: torch.Tensor, b_local: torch.Tensor, y0_local: torch.Tensor | None = None):
        lo, hi = self.row_range()
        nl = hi - lo
        if tuple(a_local.shape) != (nl, self.n) or tuple(b_local.shape) != (nl,):
            raise ValueError(f"rank {self.rank}: expected A_local {(nl, self.n)} and b_local {(nl,)}")
        a_local = a_local.to(self.dtype).contiguous()
        b_local = b_local.to(self.dtype).contiguous()
        flags = 0 if self.max_steps is None else nat.MAXSTEPS_GIVEN
        ms = 10 * self.n if self.max_steps is None else int(self.max_steps)
        restart = min(self.restart, self.n)
        if y0_local is not None:
            x = y0_local.to(self.dtype).contiguous().clone()
            flags |= nat.HAS_Y0
        else:
            x = torch.empty(nl, dtype=self.dtype, device=self.device)
        result = torch.empty(1, dtype=torch.int32, device=self.device)
        steps = torch.empty(1, dtype=torch.int32, device=self.device)
        ws_bytes = nat.fn(f"lxb_gmres_rowsharded_workspace_{self.sfx}")(nl, restart)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        nat.call(f"lxb_gmres_rowsharded_{self.sfx}", a_local.data_ptr(), b_local.data_ptr(), x.data_ptr(),
                 result.data_ptr(), steps.data_ptr(), self.n, nl, lo, self.rtol, self.atol, ms, restart,
                 self.stagnation_iters, flags, ws.data_ptr(), ws_bytes, self.peers_dev, self.world, self.rank,
                 torch.cuda.current_stream().cuda_stream)
        return x, result[0], steps[0]


class RowShardedCG(RowShardedGMRES):
    """CG (lineax/_solver/cg.py semantics, no preconditioner) on a row-partitioned SPD operator; same calling
    convention as `RowShardedGMRES` (csrc/cg_dist.cu)."""

    def __init__(self, n: int, rtol: float, atol: float, *, stabilise_every=10, max_steps=None, is_nsd=False,
                 dtype=torch.float32, group=None):
        super().__init__(n, rtol, atol, max_steps=max_steps, dtype=dtype, group=group)
        self.stabilise_every, self.is_nsd = stabilise_every, bool(is_nsd)

    def solve(self, a_local: torch.Tensor, b_local: torch.Tensor, y0_local: torch.Tensor | None = None):
        lo, hi = self.row_range()
        nl = hi - lo
        if tuple(a_local.shape) != (nl, self.n) or tuple(b_local.shape) != (nl,):
            raise ValueError(f"rank {self.rank}: expected A_local {(nl, self.n)} and b_local {(nl,)}")
        a_local = a_local.to(self.dtype).contiguous()
        b_local = b_local.to(self.dtype).contiguous()
        flags = (0 if self.max_steps is None else nat.MAXSTEPS_GIVEN) | (nat.NSD if self.is_nsd else 0)
        ms = 10 * self.n if self.max_steps is None else int(self.max_steps)
        if y0_local is not None:
            x = y0_local.to(self.dtype).contiguous().clone()
            flags |= nat.HAS_Y0
        else:
            x = torch.empty(nl, dtype=self.dtype, device=self.device)
        result = torch.empty(1, dtype=torch.int32, device=self.device)
        steps = torch.empty(1, dtype=torch.int32, device=self.device)
        ws_bytes = nat.fn(f"lxb_cg_rowsharded_workspace_{self.sfx}")(nl)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        se = 0 if self.stabilise_every is None else int(self.stabilise_every)
        nat.call(f"lxb_cg_rowsharded_{self.sfx}", a_local.data_ptr(), b_local.data_ptr(), x.data_ptr(),
                 result.data_ptr(), steps.data_ptr(), self.n, nl, lo, self.rtol, self.atol, ms, se, flags,
                 ws.data_ptr(), ws_bytes, self.peers_dev, self.world, self.rank,
                 torch.cuda.current_stream().cuda_stream)
        return x, result[0], steps[0]


class RowShardedBiCGStab(RowShardedGMRES):
    """BiCGStab (lineax/_solver/bicgstab.py semantics, no preconditioner) on a row-partitioned operator."""

    def __init__(self, n: int, rtol: float, atol: float, *, max_steps=None, x64=None, dtype=torch.float32, group=None):
        super().__init__(n, rtol, atol, max_steps=max_steps, dtype=dtype, group=group)
        self.x64 = (dtype == torch.float64) if x64 is None else bool(x64)

    def solve(self, a_local: torch.Tensor, b_local: torch.Tensor, y0_local: torch.Tensor | None = None):
        lo, hi = self.row_range()
        nl = hi - lo
        if tuple(a_local.shape) != (nl, self.n) or tuple(b_local.shape) != (nl,):
            raise ValueError(f"rank {self.rank}: expected A_local {(nl, self.n)} and b_local {(nl,)}")
        a_local = a_local.to(self.dtype).contiguous()
        b_local = b_local.to(self.dtype).contiguous()
        flags = (0 if self.max_steps is None else nat.MAXSTEPS_GIVEN) | (nat.X64_BREAKDOWN if self.x64 else 0)
        ms = 10 * self.n if self.max_steps is None else int(self.max_steps)
        if y0_local is not None:
            x = y0_local.to(self.dtype).contiguous().clone()
            flags |= nat.HAS_Y0
        else:
            x = torch.empty(nl, dtype=self.dtype, device=self.device)
        result = torch.empty(1, dtype=torch.int32, device=self.device)
        steps = torch.empty(1, dtype=torch.int32, device=self.device)
        ws_bytes = nat.fn(f"lxb_cg_rowsharded_workspace_{self.sfx}")(nl)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        nat.call(f"lxb_bicgstab_rowsharded_{self.sfx}", a_local.data_ptr(), b_local.data_ptr(), x.data_ptr(),
                 result.data_ptr(), steps.data_ptr(), self.n, nl, lo, self.rtol, self.atol, ms, flags,
                 ws.data_ptr(), ws_bytes, self.peers_dev, self.world, self.rank,
                 torch.cuda.current_stream().cuda_stream)
        return x, result[0], steps[0]


class RowShardedLSMR:
    """LSMR (lineax/_solver/lsmr.py semantics) on a tall dense operator partitioned by ROWS.

    Every rank constructs it with the same arguments and then calls `solve(A_local, b_local)` with
    its contiguous block of rows (`row_range(rank)`); returns `(x, result, num_steps, stats)` where
    the length-`n` solution and the statistics (same keys as `lineax_b200.LSMR`) are replicated
    bit-identically on every rank.
    """

    def __init__(self, m: int, n: int, rtol: float, atol: float, *, conlim: float = 1e8, max_steps=None,
                 dtype=torch.float32, group=None):
        import torch.distributed._symmetric_memory as symm

        self.m, self.n, self.rtol, self.atol = int(m), int(n), float(rtol), float(atol)
        self.conlim, self.max_steps = float(conlim), max_steps
        self.dtype, self.group = dtype, group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.bounds = shard_bounds(self.m, self.world)
        self.sfx = nat.suffix(dtype)
        self.device = torch.device("cuda", torch.cuda.current_device())
        nbytes = nat.fn(f"lxb_lsmr_rowsharded_symm_bytes_{self.sfx}")(self.n, self.world)
        self.symm_buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.symm_buf.zero_()
        self.handle = symm.rendezvous(self.symm_buf, self.group)
        self.peers_dev = int(self.handle.buffer_ptrs_dev)
        torch.cuda.synchronize()
        dist.barrier(self.group)  # every rank's flags are zero before the first kernel starts

    def row_range(self, rank=None):
        r = self.rank if rank is None else rank
        return self.bounds[r], self.bounds[r + 1]

    def solve(self, a_local: torch.Tensor, b_local: torch.Tensor, y0: torch.Tensor | None = None):
        lo, hi = self.row_range()
        ml = hi - lo
        if tuple(a_local.shape) != (ml, self.n) or tuple(b_local.shape) != (ml,):
            raise ValueError(f"rank {self.rank}: expected A_local {(ml, self.n)} and b_local {(ml,)}")
        a_local = a_local.to(self.dtype).contiguous()
        b_local = b_local.to(self.dtype).contiguous()
        min_dim = min(self.m, self.n)
        flags = 0
        if self.max_steps is None:  # lsmr.py:121-129, with the integer-overflow guard
            imax = torch.iinfo(torch.int32 if self.dtype == torch.float32 else torch.int64).max
            ms = imax if min_dim > imax / 10 else min_dim * 10
        else:
            ms, flags = int(self.max_steps), nat.MAXSTEPS_GIVEN
        if y0 is not None:
            x = y0.to(self.dtype).contiguous().clone()
            flags |= nat.HAS_Y0
        else:
            x = torch.empty(self.n, dtype=self.dtype, device=self.device)
        result = torch.empty(1, dtype=torch.int32, device=self.device)
        steps = torch.empty(1, dtype=torch.int32, device=self.device)
        st = torch.empty(8, dtype=self.dtype, device=self.device)
        ws_bytes = nat.fn(f"lxb_lsmr_rowsharded_workspace_{self.sfx}")(ml, self.n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        nat.call(f"lxb_lsmr_rowsharded_{self.sfx}", a_local.data_ptr(), b_local.data_ptr(), x.data_ptr(),
                 result.data_ptr(), steps.data_ptr(), st.data_ptr(), self.m, ml, self.n, self.rtol, self.atol,
                 self.conlim, ms, flags, ws.data_ptr(), ws_bytes, self.peers_dev, self.world, self.rank,
                 torch.cuda.current_stream().cuda_stream)
        stats = {"num_steps": steps[0], "istop": st[0].to(torch.int32), "norm_r": st[1], "norm_Ar": st[2],
                 "norm_A": st[3], "cond_A": st[4], "norm_x": st[5]}
        return x, result[0], steps[0], stats


class RowShardedQR:
    """QR least squares (lineax/_solver/qr.py:55-94, tall branch) on a tall dense operator partitioned by
    ROWS: communication-avoiding TSQR.  Every rank factors its own block with the blocked Householder
    kernels (csrc/qr_large.cu) and reduces its right-hand side to `(Q_p^T b_p)[:n]`; ONE all-gather stacks
    the P upper-triangular `R_p` (n x n) and reduced right-hand sides, and every rank finishes with the
    same small QR of the stacked `P n x n` matrix -- the solution is replicated bit-identically.
    Requires every rank's block to be tall (`m_local >= n`)."""

    def __init__(self, m: int, n: int, *, dtype=torch.float32, group=None, device=None):
        self.m, self.n, self.dtype = int(m), int(n), dtype
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.bounds = shard_bounds(self.m, self.world)
        if min(self.bounds[r + 1] - self.bounds[r] for r in range(self.world)) < self.n:
            raise ValueError("row-sharded QR needs m_local >= n on every rank")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def row_range(self, rank=None):
        r = self.rank if rank is None else rank
        return self.bounds[r], self.bounds[r + 1]

    def solve(self, a_local: torch.Tensor, b_local: torch.Tensor) -> torch.Tensor:
        lo, hi = self.row_range()
        if tuple(a_local.shape) != (hi - lo, self.n) or tuple(b_local.shape) != (hi - lo,):
            raise ValueError(f"rank {self.rank}: expected A_local {(hi - lo, self.n)} and b_local {(hi - lo,)}")
        n = self.n
        aq, taus = _ops.qr_factor(a_local.to(self.dtype))
        if self.world == 1:
            return _ops.qr_solve(aq, taus, b_local.to(self.dtype), False)
        c = _ops.qr_apply_qt(aq, taus, b_local.to(self.dtype))
        r = torch.triu(aq[:n])  # (masking only: the Householder vectors below the diagonal are dropped)
        del aq
        stack = torch.empty(self.world * n, n, dtype=self.dtype, device=self.device)
        cstack = torch.empty(self.world * n, dtype=self.dtype, device=self.device)
        dist.all_gather_into_tensor(stack, r.contiguous(), group=self.group)
        dist.all_gather_into_tensor(cstack, c.contiguous(), group=self.group)
        aq2, taus2 = _ops.qr_factor(stack)
        return _ops.qr_solve(aq2, taus2, cstack, False)


class RowShardedMatrixLinearOperator(AbstractLinearOperator):
    """A dense `rows x cols` operator of which this rank holds the contiguous block of rows
    `shard_bounds(rows, world)[rank : rank + 2]`.  `lx.linear_solve(op, b_local, solver)` with
    `lx.GMRES` (square: `b_local` / the solution are the local row slices), `lx.LSMR` or `lx.QR`
    (tall: `b_local` is the local slice, the solution is replicated) runs the row-sharded kernels
    (lineax/_solve.py:656-806 is the entry point they sit behind)."""

    def __init__(self, local_rows: torch.Tensor, rows: int, group=None):
        self.local = local_rows
        self.rows, self.cols = int(rows), int(local_rows.shape[-1])
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.bounds = shard_bounds(self.rows, self.world)
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        if local_rows.shape[0] != hi - lo:
            raise ValueError(f"rank {self.rank} must hold rows [{lo}, {hi}) of the operator")
        self._solvers = {}

    # the operator interface speaks about the LOCAL block (what this process can address)
    def mv(self, vector):
        return _ops.matvec(self.local, vector, False)

    def as_matrix(self):
        raise ValueError("a row-sharded operator is never materialised on one device")

    def transpose(self):
        raise NotImplementedError("transposes of row-sharded operators are not supported")

    def in_structure(self):
        # square systems: every vector is sharded like the rows; tall systems: x is replicated
        n = self.local.shape[0] if self.rows == self.cols else self.cols
        return ShapeDtypeStruct((n,), self.local.dtype)

    def out_structure(self):
        return ShapeDtypeStruct((self.local.shape[0],), self.local.dtype)

    def sharded_solver(self, kind: str, key: tuple, make):
        """Symmetric-memory buffers are allocated once per (solver kind, parameters) and reused."""
        k = (kind,) + key
        if k not in self._solvers:
            self._solvers[k] = make()
        return self._solvers[k]
