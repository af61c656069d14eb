"""Tensor-level entry points: thin torch plumbing over the C ABI.

Every function accepts arbitrary leading batch dimensions (broadcast between
operands) and consumes them natively: the whole batch is ONE kernel launch.
Each is registered as a `torch.library` custom op with a vmap rule, which is the
torch analogue of the `jax.ffi.ffi_call(..., vmap_method=...)` batching rule the
north star asks for: `torch.func.vmap(lambda A, b: linear_solve(MatrixLinearOperator(A), b, LU()).value)`
lowers to a single batched launch (reference: the auto-vmap rule of
`eqxi.create_vprim`, lineax/_solve.py:320-327).

PyTorch is used only for device memory, streams and dispatch; no arithmetic of the
solve is done by torch.
"""
import math
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native as nat


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _batch(t: Tensor, core: int) -> Tuple[int, ...]:
    return tuple(t.shape[: t.ndim - core])


def _full_batch(*pairs) -> Tuple[int, ...]:
    return tuple(torch.broadcast_shapes(*[_batch(t, c) for t, c in pairs if t is not None]))


def _operand(t: Tensor, core: int, full: Tuple[int, ...]):
    """-> (contiguous tensor, element stride between systems; 0 = broadcast)."""
    bs = _batch(t, core)
    core_shape = tuple(t.shape[t.ndim - core:])
    per = math.prod(core_shape)
    if all(s == 1 for s in bs) and math.prod(full) != 1:
        return t.reshape(core_shape).contiguous(), 0
    if bs != full:
        t = t.expand(full + core_shape)
    return t.contiguous(), per


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _workspace(name: str, device, *args):
    nbytes = nat.fn(name)(*args)
    if not nbytes:
        return None, 0
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "lineax_b200 kernels run on CUDA (sm_100a) only and have no CPU fallback; "
                f"got a tensor on {t.device}."
            )


def _define(name: str, fn, *, n_out: int, device_types="cuda"):
    op = torch.library.custom_op(f"lineax_b200::{name}", fn, mutates_args=(), device_types=device_types)

    def rule(info, in_dims, *args):
        new = []
        for a, d in zip(args, in_dims):
            if isinstance(a, Tensor):
                new.append(a.movedim(d, 0) if d is not None else a.unsqueeze(0))
            else:
                new.append(a)
        out = op(*new)
        if n_out == 1:
            return out, 0
        return tuple(out), tuple(0 for _ in range(n_out))

    torch.library.register_vmap(op, rule)
    return op


# --------------------------------------------------------------------- LU ----
def _use_multi(stride_factor: int, stride_b: int, nvec: int, n: int, dtype) -> bool:
    """ONE factorisation against many vectors (vmap(in_axes=(None, 0)), `state=` reuse, `invert`): the
    thread-per-vector multi-RHS kernels (csrc/multi_rhs.cu) read the factor once per tile of vectors."""
    return stride_factor == 0 and stride_b == n and nvec >= 8 and n * 33 * dtype.itemsize <= 200 * 1024


def _lu_factor(a: Tensor) -> Tuple[Tensor, Tensor]:
    _check_cuda(a)
    n = a.shape[-1]
    full = _batch(a, 2)
    B = math.prod(full)
    sfx = nat.suffix(a.dtype)
    with torch.cuda.device(a.device):
        a_ = a.contiguous()
        lu = torch.empty_like(a_)
        piv = torch.empty(full + (n,), dtype=torch.int32, device=a.device)
        nat.call(f"lxb_lu_factor_{sfx}", a_.data_ptr(), n * n, lu.data_ptr(), piv.data_ptr(), B, n,
                 _stream())
    return lu, piv


def _lu_solve(lu: Tensor, piv: Tensor, b: Tensor, trans: bool) -> Tensor:
    _check_cuda(lu, piv, b)
    n = lu.shape[-1]
    full = _full_batch((lu, 2), (piv, 1), (b, 1))
    B = math.prod(full)
    sfx = nat.suffix(lu.dtype)
    with torch.cuda.device(lu.device):
        lu_, s_lu = _operand(lu, 2, full)
        piv_, s_p = _operand(piv, 1, full)
        b_, s_b = _operand(b.to(lu.dtype), 1, full)
        x = torch.empty(full + (n,), dtype=lu.dtype, device=lu.device)
        if _use_multi(s_lu, s_b, B, n, lu.dtype) and s_p == 0:
            nat.call(f"lxb_lu_solve_multi_{sfx}", lu_.data_ptr(), n * n, piv_.data_ptr(), n, b_.data_ptr(),
                     x.data_ptr(), 1, n, B, nat.TRANS if trans else 0, _stream())
            return x
        nat.call(f"lxb_lu_solve_{sfx}", lu_.data_ptr(), s_lu, piv_.data_ptr(), s_p, b_.data_ptr(), s_b,
                 x.data_ptr(), B, n, nat.TRANS if trans else 0, _stream())
    return x


# lu buffer doubles as workspace for systems too large for shared memory (lu.cu, Tier M)
_LU_SMEM_LIMIT = {torch.float32: 238, torch.float64: 168}


def _lu_factor_solve(a: Tensor, b: Tensor, keep_state: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _check_cuda(a, b)
    n = a.shape[-1]
    full = _full_batch((a, 2), (b, 1))
    B = math.prod(full)
    sfx = nat.suffix(a.dtype)
    need_lu = keep_state or n > _LU_SMEM_LIMIT[a.dtype]
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a, 2, full)
        b_, s_b = _operand(b.to(a.dtype), 1, full)
        x = torch.empty(full + (n,), dtype=a.dtype, device=a.device)
        if need_lu:
            lu = torch.empty(full + (n, n), dtype=a.dtype, device=a.device)
            piv = torch.empty(full + (n,), dtype=torch.int32, device=a.device)
        else:
            lu = torch.empty(full + (0, 0), dtype=a.dtype, device=a.device)
            piv = torch.empty(full + (0,), dtype=torch.int32, device=a.device)
        nat.call(f"lxb_lu_factor_solve_{sfx}", a_.data_ptr(), s_a, b_.data_ptr(), s_b, x.data_ptr(),
                 lu.data_ptr() if need_lu else None, piv.data_ptr() if need_lu else None, B, n,
                 _stream())
    return x, lu, piv


lu_factor = _define("lu_factor", _lu_factor, n_out=2)
lu_solve = _define("lu_solve", _lu_solve, n_out=1)
lu_factor_solve = _define("lu_factor_solve", _lu_factor_solve, n_out=3)


# --------------------------------------------------------------------- CG ----
def _cg(a: Tensor, b: Tensor, precond: Optional[Tensor], y0: Optional[Tensor], rtol: float,
        atol: float, max_steps: int, stabilise_every: int, flags: int) -> Tuple[Tensor, Tensor, Tensor]:
    _check_cuda(a, b, precond, y0)
    n = a.shape[-1]
    full = _full_batch((a, 2), (b, 1), (precond, 2), (y0, 1))
    B = math.prod(full)
    sfx = nat.suffix(a.dtype)
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a, 2, full)
        b_, s_b = _operand(b.to(a.dtype), 1, full)
        m_, s_m = (None, 0) if precond is None else _operand(precond.to(a.dtype), 2, full)
        if y0 is not None:
            x = y0.to(a.dtype).expand(full + (n,)).contiguous().clone()
            flags |= nat.HAS_Y0
        else:
            x = torch.empty(full + (n,), dtype=a.dtype, device=a.device)
        result = torch.empty(full, dtype=torch.int32, device=a.device)
        steps = torch.empty(full, dtype=torch.int32, device=a.device)
        ws, ws_bytes = _workspace(f"lxb_cg_workspace_{sfx}", a.device, B, n)
        nat.call(f"lxb_cg_{sfx}", a_.data_ptr(), s_a, b_.data_ptr(), s_b, _ptr(m_), s_m, x.data_ptr(),
                 result.data_ptr(), steps.data_ptr(), B, n, rtol, atol, max_steps, stabilise_every,
                 flags, _ptr(ws), ws_bytes, _stream())
    return x, result, steps


cg = _define("cg", _cg, n_out=3)


# ---------------------------------------------------------- post-processing ----
def _postprocess(x: Tensor, b: Tensor, result: Tensor) -> Tensor:
    """lineax/_solve.py:104-123 on flattened solution / vector (last dim = elements)."""
    _check_cuda(x, b, result)
    full = _full_batch((x, 1), (b, 1), (result, 0))
    B = math.prod(full)
    dt = torch.promote_types(x.dtype, b.dtype)
    sfx = nat.suffix(dt)
    with torch.cuda.device(x.device):
        x_, s_x = _operand(x.to(dt), 1, full)
        b_, s_b = _operand(b.to(dt), 1, full)
        out = result.to(torch.int32).expand(full).contiguous().clone()
        nat.call(f"lxb_postprocess_{sfx}", x_.data_ptr(), s_x, x.shape[-1], b_.data_ptr(), s_b,
                 b.shape[-1], out.data_ptr(), B, _stream())
    return out


postprocess = _define("postprocess", _postprocess, n_out=1)


def _throw_if_failed(result: Tensor) -> Tensor:
    """`throw=True` (lineax/_solve.py:124-128): raise on any non-successful code (host sync)."""
    from ._solution import RESULTS, LinearSolveError

    flat = result.reshape(-1)
    # one scalar leaves the device (the reference's error_if is asynchronous; callers that issue many
    # small solves should pass throw=False and inspect `Solution.result` themselves)
    if bool((flat != 0).any()):
        codes = flat.tolist()
        bad = [c for c in codes if c != 0]
        where = "" if len(codes) == 1 else f" ({len(bad)} of {len(codes)} systems failed)"
        raise LinearSolveError(RESULTS[bad[0]] + where)
    return result.clone()


throw_if_failed = _define("throw_if_failed", _throw_if_failed, n_out=1, device_types=None)  # host logic only


# ------------------------------------------------------ operator application ----
def _matvec(a: Tensor, x: Tensor, trans: bool) -> Tensor:
    _check_cuda(a, x)
    m, n = a.shape[-2], a.shape[-1]
    dt = torch.promote_types(a.dtype, x.dtype)
    full = _full_batch((a, 2), (x, 1))
    B = math.prod(full)
    sfx = nat.suffix(dt)
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a.to(dt), 2, full)
        x_, s_x = _operand(x.to(dt), 1, full)
        y = torch.empty(full + ((n if trans else m),), dtype=dt, device=a.device)
        nat.call(f"lxb_matvec_{sfx}", a_.data_ptr(), s_a, x_.data_ptr(), s_x, y.data_ptr(), B, m, n,
                 nat.TRANS if trans else 0, _stream())
    return y


matvec = _define("matvec", _matvec, n_out=1)


def _gram(a: Tensor, aat: bool) -> Tensor:
    """A^T A (aat False) or A A^T (aat True): the operator of the normal equations (normal.py:111-117)."""
    _check_cuda(a)
    m, n = a.shape[-2], a.shape[-1]
    full = _batch(a, 2)
    B = math.prod(full)
    sfx = nat.suffix(a.dtype)
    g = m if aat else n
    with torch.cuda.device(a.device):
        a_ = a.contiguous()
        out = torch.empty(full + (g, g), dtype=a.dtype, device=a.device)
        nat.call(f"lxb_gram_{sfx}", a_.data_ptr(), m * n, out.data_ptr(), B, m, n, nat.TRANS if aat else 0,
                 _stream())
    return out


gram = _define("gram", _gram, n_out=1)


def _diag_mv(d: Tensor, x: Tensor) -> Tensor:
    """Elementwise product of equally-shaped leaves (no batch semantics needed)."""
    _check_cuda(d, x)
    dt = torch.promote_types(d.dtype, x.dtype)
    shape = torch.broadcast_shapes(d.shape, x.shape)
    sfx = nat.suffix(dt)
    with torch.cuda.device(d.device):
        d_ = d.to(dt).expand(shape).contiguous()
        x_ = x.to(dt).expand(shape).contiguous()
        y = torch.empty(shape, dtype=dt, device=d.device)
        n = y.numel()
        if n:
            nat.call(f"lxb_diag_mv_{sfx}", d_.data_ptr(), n, x_.data_ptr(), n, y.data_ptr(), 1, n, _stream())
    return y


diag_mv = _define("diag_mv", _diag_mv, n_out=1)


def _tridiag_mv(d: Tensor, dl: Tensor, du: Tensor, x: Tensor) -> Tensor:
    _check_cuda(d, dl, du, x)
    n = d.shape[-1]
    dt = torch.promote_types(d.dtype, x.dtype)
    full = _full_batch((d, 1), (dl, 1), (du, 1), (x, 1))
    B = math.prod(full)
    sfx = nat.suffix(dt)
    with torch.cuda.device(d.device):
        d_ = d.to(dt).expand(full + (n,)).contiguous()
        dl_ = dl.to(dt).expand(full + (max(n - 1, 0),)).contiguous()
        du_ = du.to(dt).expand(full + (max(n - 1, 0),)).contiguous()
        x_, s_x = _operand(x.to(dt), 1, full)
        y = torch.empty(full + (n,), dtype=dt, device=d.device)
        nat.call(f"lxb_tridiag_mv_{sfx}", d_.data_ptr(), dl_.data_ptr() if n > 1 else None,
                 du_.data_ptr() if n > 1 else None, n, x_.data_ptr(), s_x, y.data_ptr(), B, n, _stream())
    return y


tridiag_mv = _define("tridiag_mv", _tridiag_mv, n_out=1)


def _norms(x: Tensor) -> Tensor:
    """-> [..., 3] = (two_norm, max_norm, 0) of the last dimension."""
    _check_cuda(x)
    full = _batch(x, 1)
    sfx = nat.suffix(x.dtype)
    with torch.cuda.device(x.device):
        x_ = x.contiguous()
        out = torch.empty(full + (3,), dtype=x.dtype, device=x.device)
        nat.call(f"lxb_norms_{sfx}", x_.data_ptr(), x.shape[-1], None, 0, out.data_ptr(),
                 math.prod(full), x.shape[-1], _stream())
    return out


_norms_op = _define("norms", _norms, n_out=1)


def norms(x: Tensor):
    out = _norms_op(x)
    return out[..., 0], out[..., 1]


def _dot(x: Tensor, y: Tensor) -> Tensor:
    _check_cuda(x, y)
    dt = torch.promote_types(x.dtype, y.dtype)
    full = _full_batch((x, 1), (y, 1))
    sfx = nat.suffix(dt)
    with torch.cuda.device(x.device):
        x_, s_x = _operand(x.to(dt), 1, full)
        y_, s_y = _operand(y.to(dt), 1, full)
        out = torch.empty(full + (3,), dtype=dt, device=x.device)
        nat.call(f"lxb_norms_{sfx}", x_.data_ptr(), s_x, y_.data_ptr(), s_y, out.data_ptr(),
                 math.prod(full), x.shape[-1], _stream())
    return out[..., 2].contiguous()


dot = _define("dot", _dot, n_out=1)


# ------------------------------------------------------------ other Krylov ----
def _krylov_common(a, b, precond, y0, core_a=2):
    _check_cuda(a, b, precond, y0)
    full = _full_batch((a, core_a), (b, 1), (precond, 2), (y0, 1))
    return full, math.prod(full), nat.suffix(a.dtype)


def _bicgstab(a: Tensor, b: Tensor, precond: Optional[Tensor], y0: Optional[Tensor], rtol: float,
              atol: float, max_steps: int, flags: int) -> Tuple[Tensor, Tensor, Tensor]:
    n = a.shape[-1]
    full, B, sfx = _krylov_common(a, b, precond, y0)
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a, 2, full)
        b_, s_b = _operand(b.to(a.dtype), 1, full)
        m_, s_m = (None, 0) if precond is None else _operand(precond.to(a.dtype), 2, full)
        if y0 is not None:
            x = y0.to(a.dtype).expand(full + (n,)).contiguous().clone()
            flags |= nat.HAS_Y0
        else:
            x = torch.empty(full + (n,), dtype=a.dtype, device=a.device)
        result = torch.empty(full, dtype=torch.int32, device=a.device)
        steps = torch.empty(full, dtype=torch.int32, device=a.device)
        ws, ws_bytes = _workspace(f"lxb_bicgstab_workspace_{sfx}", a.device, B, n)
        nat.call(f"lxb_bicgstab_{sfx}", a_.data_ptr(), s_a, b_.data_ptr(), s_b, _ptr(m_), s_m,
                 x.data_ptr(), result.data_ptr(), steps.data_ptr(), B, n, rtol, atol, max_steps, flags,
                 _ptr(ws), ws_bytes, _stream())
    return x, result, steps


bicgstab = _define("bicgstab", _bicgstab, n_out=3)


def _gmres(a: Tensor, b: Tensor, precond: Optional[Tensor], y0: Optional[Tensor], rtol: float,
           atol: float, max_steps: int, restart: int, stagnation_iters: int,
           flags: int) -> Tuple[Tensor, Tensor, Tensor]:
    n = a.shape[-1]
    full, B, sfx = _krylov_common(a, b, precond, y0)
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a, 2, full)
        b_, s_b = _operand(b.to(a.dtype), 1, full)
        m_, s_m = (None, 0) if precond is None else _operand(precond.to(a.dtype), 2, full)
        if y0 is not None:
            x = y0.to(a.dtype).expand(full + (n,)).contiguous().clone()
            flags |= nat.HAS_Y0
        else:
            x = torch.empty(full + (n,), dtype=a.dtype, device=a.device)
        result = torch.empty(full, dtype=torch.int32, device=a.device)
        steps = torch.empty(full, dtype=torch.int32, device=a.device)
        ws_bytes = nat.fn(f"lxb_gmres_workspace_{sfx}")(B, n, restart)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device) if ws_bytes else None
        nat.call(f"lxb_gmres_{sfx}", a_.data_ptr(), s_a, b_.data_ptr(), s_b, _ptr(m_), s_m, x.data_ptr(),
                 result.data_ptr(), steps.data_ptr(), B, n, rtol, atol, max_steps, restart,
                 stagnation_iters, flags, _ptr(ws), ws_bytes, _stream())
    return x, result, steps


gmres = _define("gmres", _gmres, n_out=3)


def _lsmr(a: Tensor, b: Tensor, y0: Optional[Tensor], rtol: float, atol: float, conlim: float,
          max_steps: int, flags: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    _check_cuda(a, b, y0)
    m, n = a.shape[-2], a.shape[-1]
    full = _full_batch((a, 2), (b, 1), (y0, 1))
    B = math.prod(full)
    sfx = nat.suffix(a.dtype)
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a, 2, full)
        b_, s_b = _operand(b.to(a.dtype), 1, full)
        if y0 is not None:
            x = y0.to(a.dtype).expand(full + (n,)).contiguous().clone()
            flags |= nat.HAS_Y0
        else:
            x = torch.empty(full + (n,), dtype=a.dtype, device=a.device)
        result = torch.empty(full, dtype=torch.int32, device=a.device)
        steps = torch.empty(full, dtype=torch.int32, device=a.device)
        stats = torch.empty(full + (8,), dtype=a.dtype, device=a.device)
        ws, ws_bytes = _workspace(f"lxb_lsmr_workspace_{sfx}", a.device, B, m, n)
        nat.call(f"lxb_lsmr_{sfx}", a_.data_ptr(), s_a, b_.data_ptr(), s_b, x.data_ptr(),
                 result.data_ptr(), steps.data_ptr(), stats.data_ptr(), B, m, n, rtol, atol, conlim,
                 max_steps, flags, _ptr(ws), ws_bytes, _stream())
    return x, result, steps, stats


lsmr = _define("lsmr", _lsmr, n_out=4)


# ------------------------------------------------------------ other direct ----
def _cholesky_factor(a: Tensor, nsd: bool) -> Tensor:
    _check_cuda(a)
    n = a.shape[-1]
    full = _batch(a, 2)
    sfx = nat.suffix(a.dtype)
    with torch.cuda.device(a.device):
        a_ = a.contiguous()
        f = torch.empty_like(a_)
        nat.call(f"lxb_cholesky_factor_{sfx}", a_.data_ptr(), n * n, f.data_ptr(), math.prod(full), n,
                 nat.NSD if nsd else 0, _stream())
    return f


def _cholesky_solve(f: Tensor, b: Tensor, nsd: bool) -> Tensor:
    _check_cuda(f, b)
    n = f.shape[-1]
    full = _full_batch((f, 2), (b, 1))
    sfx = nat.suffix(f.dtype)
    with torch.cuda.device(f.device):
        f_, s_f = _operand(f, 2, full)
        b_, s_b = _operand(b.to(f.dtype), 1, full)
        x = torch.empty(full + (n,), dtype=f.dtype, device=f.device)
        if _use_multi(s_f, s_b, math.prod(full), n, f.dtype):
            nat.call(f"lxb_cholesky_solve_multi_{sfx}", f_.data_ptr(), n * n, b_.data_ptr(), x.data_ptr(), 1, n,
                     math.prod(full), nat.NSD if nsd else 0, _stream())
            return x
        nat.call(f"lxb_cholesky_solve_{sfx}", f_.data_ptr(), s_f, b_.data_ptr(), s_b, x.data_ptr(),
                 math.prod(full), n, nat.NSD if nsd else 0, _stream())
    return x


cholesky_factor = _define("cholesky_factor", _cholesky_factor, n_out=1)
cholesky_solve = _define("cholesky_solve", _cholesky_solve, n_out=1)


def _qr_factor(a: Tensor) -> Tuple[Tensor, Tensor]:
    _check_cuda(a)
    m, n = a.shape[-2], a.shape[-1]
    rows, cols = max(m, n), min(m, n)
    full = _batch(a, 2)
    sfx = nat.suffix(a.dtype)
    with torch.cuda.device(a.device):
        a_ = a.contiguous()
        out = torch.empty(full + (rows, cols), dtype=a.dtype, device=a.device)
        taus = torch.empty(full + (cols,), dtype=a.dtype, device=a.device)
        ws, ws_bytes = _workspace(f"lxb_qr_factor_workspace_{sfx}", a.device, math.prod(full), m, n)
        nat.call(f"lxb_qr_factor_{sfx}", a_.data_ptr(), m * n, out.data_ptr(), taus.data_ptr(),
                 math.prod(full), m, n, _ptr(ws), ws_bytes, _stream())
    return out, taus


def _qr_apply_qt(a: Tensor, taus: Tensor, b: Tensor) -> Tensor:
    """(Q^T b)[:cols] for tall factors (qr.py:89-90 without the trailing triangular solve)."""
    _check_cuda(a, taus, b)
    rows, cols = a.shape[-2], a.shape[-1]
    full = _full_batch((a, 2), (taus, 1), (b, 1))
    sfx = nat.suffix(a.dtype)
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a, 2, full)
        t_, s_t = _operand(taus, 1, full)
        b_, s_b = _operand(b.to(a.dtype), 1, full)
        y = torch.empty(full + (cols,), dtype=a.dtype, device=a.device)
        ws, ws_bytes = _workspace(f"lxb_qr_solve_workspace_{sfx}", a.device, math.prod(full), rows, cols)
        nat.call(f"lxb_qr_solve_{sfx}", a_.data_ptr(), s_a, t_.data_ptr(), s_t, b_.data_ptr(), s_b,
                 y.data_ptr(), math.prod(full), rows, cols, nat.QT_ONLY, _ptr(ws), ws_bytes, _stream())
    return y


qr_apply_qt = _define("qr_apply_qt", _qr_apply_qt, n_out=1)


def _qr_solve(a: Tensor, taus: Tensor, b: Tensor, trans: bool) -> Tensor:
    _check_cuda(a, taus, b)
    rows, cols = a.shape[-2], a.shape[-1]
    full = _full_batch((a, 2), (taus, 1), (b, 1))
    sfx = nat.suffix(a.dtype)
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a, 2, full)
        t_, s_t = _operand(taus, 1, full)
        b_, s_b = _operand(b.to(a.dtype), 1, full)
        x = torch.empty(full + ((rows if trans else cols),), dtype=a.dtype, device=a.device)
        ws, ws_bytes = _workspace(f"lxb_qr_solve_workspace_{sfx}", a.device, math.prod(full), rows, cols)
        nat.call(f"lxb_qr_solve_{sfx}", a_.data_ptr(), s_a, t_.data_ptr(), s_t, b_.data_ptr(), s_b,
                 x.data_ptr(), math.prod(full), rows, cols, nat.TRANS if trans else 0, _ptr(ws), ws_bytes,
                 _stream())
    return x


qr_factor = _define("qr_factor", _qr_factor, n_out=2)
qr_solve = _define("qr_solve", _qr_solve, n_out=1)


def _tridiagonal_solve(d: Tensor, dl: Tensor, du: Tensor, b: Tensor) -> Tensor:
    _check_cuda(d, dl, du, b)
    n = d.shape[-1]
    full = _full_batch((d, 1), (dl, 1), (du, 1), (b, 1))
    B = math.prod(full)
    sfx = nat.suffix(d.dtype)
    with torch.cuda.device(d.device):
        d_ = d.expand(full + (n,)).contiguous()
        dl_ = dl.to(d.dtype).expand(full + (max(n - 1, 0),)).contiguous()
        du_ = du.to(d.dtype).expand(full + (max(n - 1, 0),)).contiguous()
        b_, s_b = _operand(b.to(d.dtype), 1, full)
        x = torch.empty(full + (n,), dtype=d.dtype, device=d.device)
        ws_bytes = nat.fn(f"lxb_tridiagonal_workspace_{sfx}")(B, n)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=d.device)
        nat.call(f"lxb_tridiagonal_solve_{sfx}", d_.data_ptr(), dl_.data_ptr() if n > 1 else None,
                 du_.data_ptr() if n > 1 else None, n, b_.data_ptr(), s_b, x.data_ptr(), B, n,
                 ws.data_ptr(), ws_bytes, _stream())
    return x


tridiagonal_solve = _define("tridiagonal_solve", _tridiagonal_solve, n_out=1)


def _diagonal_solve(d: Tensor, b: Tensor, rcond: float) -> Tensor:
    _check_cuda(d, b)
    n = d.shape[-1]
    full = _full_batch((d, 1), (b, 1))
    dt = torch.promote_types(d.dtype, b.dtype)
    sfx = nat.suffix(dt)
    with torch.cuda.device(d.device):
        d_, s_d = _operand(d.to(dt), 1, full)
        b_, s_b = _operand(b.to(dt), 1, full)
        x = torch.empty(full + (n,), dtype=dt, device=d.device)
        nat.call(f"lxb_diagonal_solve_{sfx}", d_.data_ptr(), s_d, b_.data_ptr(), s_b, x.data_ptr(),
                 math.prod(full), n, rcond, _stream())
    return x


diagonal_solve = _define("diagonal_solve", _diagonal_solve, n_out=1)


def _triangular_solve(a: Tensor, b: Tensor, lower: bool, unit: bool, trans: bool) -> Tensor:
    _check_cuda(a, b)
    n = a.shape[-1]
    full = _full_batch((a, 2), (b, 1))
    dt = torch.promote_types(a.dtype, b.dtype)
    sfx = nat.suffix(dt)
    flags = (nat.LOWER if lower else 0) | (nat.UNIT_DIAG if unit else 0) | (nat.TRANS if trans else 0)
    with torch.cuda.device(a.device):
        a_, s_a = _operand(a.to(dt), 2, full)
        b_, s_b = _operand(b.to(dt), 1, full)
        x = torch.empty(full + (n,), dtype=dt, device=a.device)
        if _use_multi(s_a, s_b, math.prod(full), n, dt):
            nat.call(f"lxb_triangular_solve_multi_{sfx}", a_.data_ptr(), n * n, b_.data_ptr(), x.data_ptr(), 1, n,
                     math.prod(full), flags, _stream())
            return x
        nat.call(f"lxb_triangular_solve_{sfx}", a_.data_ptr(), s_a, b_.data_ptr(), s_b, x.data_ptr(),
                 math.prod(full), n, flags, _stream())
    return x


triangular_solve = _define("triangular_solve", _triangular_solve, n_out=1)
