"""Multi-GPU path on CPU: world_size-2 `gloo` processes exercise the batch-sharding plumbing
(lineax_b200/_shard.py): each rank solves its contiguous block, results are gathered, and the
gathered answer equals the single-process one.  Numerical ops are the oracle-backed CPU doubles."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lineax_b200 as lx
    from lineax_b200._shard import local_slice, shard_bounds, solve_sharded
    from tests import cpu_kernels

    cpu_kernels.install()
    lx.set_default_device("cpu")
    rng = np.random.default_rng(0)  # same inputs on every rank
    mats = torch.as_tensor(rng.standard_normal((batch, 6, 6)) + 4 * np.eye(6))
    vecs = torch.as_tensor(rng.standard_normal((batch, 6)))
    solve = torch.func.vmap(
        lambda m, v: lx.linear_solve(lx.MatrixLinearOperator(m), v, lx.LU(), throw=False).value)
    sl = local_slice(batch)
    assert sl.stop - sl.start in (batch // world, batch // world + 1)
    x = solve_sharded(solve, mats, vecs)
    assert x.shape == (batch, 6)
    ref = np.stack([np.linalg.solve(mats[i].numpy(), vecs[i].numpy()) for i in range(batch)])
    assert np.allclose(x.numpy(), ref, atol=1e-10)
    local = solve_sharded(solve, mats, vecs, gather=False)
    assert local.shape[0] == sl.stop - sl.start
    # timing contract of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    np.save(os.path.join(out_dir, f"x{rank}.npy"), x.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 11])
def test_batch_sharding_world2_gloo(tmp_path, batch):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    x0, x1 = np.load(tmp_path / "x0.npy"), np.load(tmp_path / "x1.npy")
    assert np.array_equal(x0, x1)


def test_shard_bounds():
    from lineax_b200._shard import shard_bounds

    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(65536, 8)[-1] == 65536 and shard_bounds(3, 8)[-1] == 3


def _lsmr_rows_worker(rank, world, port, out_dir):
    """The decomposition lsmr_dist.cu uses, restated with numpy + gloo: rows of A and u sharded,
    v / x / h / hbar replicated, ONE all-reduce per iteration carrying the partial A^T u' and ||u'||^2."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import gen
    from oracle.krylov import _givens as _sym_ortho
    from lineax_b200._shard import shard_bounds

    m, n, tol = 301, 24, 1e-10
    a, b, _ = gen.tall_lstsq(3, m, n, np.float64)
    lo, hi = shard_bounds(m, world)[rank], shard_bounds(m, world)[rank + 1]
    al, u = a[lo:hi], b[lo:hi].copy()

    def allsum(vec):
        t = torch.as_tensor(np.ascontiguousarray(vec))
        dist.all_reduce(t)
        return t.numpy()

    normb = np.sqrt(allsum(np.array([u @ u]))[0])
    x = np.zeros(n)
    red = allsum(np.concatenate([al.T @ u, [u @ u]]))  # first exchange: x0 = 0 so u' = b
    beta = np.sqrt(red[n])
    u /= beta
    v = red[:n] / beta
    alpha = np.linalg.norm(v)
    v /= alpha
    h, hbar = v.copy(), np.zeros(n)
    zetabar, alphabar, rho, rhobar, cbar, sbar = alpha * beta, alpha, 1.0, 1.0, 1.0, 0.0
    betadd, betad, rhodold, tautildeold, thetatilde, zeta, delta = beta, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0
    normA2, istop, itn = alpha * alpha, 0, 0
    while istop == 0 and itn < 10 * n:
        itn += 1
        up = al @ v - alpha * u                      # local rows only
        red = allsum(np.concatenate([al.T @ up, [up @ up]]))  # THE exchange of the iteration
        beta = np.sqrt(red[n])
        u = up / beta
        v = red[:n] / beta - beta * v
        alpha = np.linalg.norm(v)
        v /= alpha
        chat, shat, alphahat = _sym_ortho(alphabar, 0.0)
        rhoold = rho
        c, s, rho = _sym_ortho(alphahat, beta)
        thetanew, alphabar = s * alpha, c * alpha
        rhobarold, zetaold, thetabar = rhobar, zeta, sbar * rho
        cbar, sbar, rhobar = _sym_ortho(cbar * rho, thetanew)
        zeta, zetabar = cbar * zetabar, -sbar * zetabar
        hbar = h - (thetabar * rho / (rhoold * rhobarold)) * hbar
        x = x + (zeta / (rho * rhobar)) * hbar
        h = v - (thetanew / rho) * h
        betaacute, betacheck = chat * betadd, -shat * betadd
        betahat, betadd = c * betaacute, -s * betaacute
        thetatildeold = thetatilde
        ctildeold, stildeold, rhotildeold = _sym_ortho(rhodold, thetabar)
        thetatilde, rhodold = stildeold * rhobar, ctildeold * rhobar
        betad = -stildeold * betad + ctildeold * betahat
        tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold
        taud = (zeta - thetatilde * tautildeold) / rhodold
        delta += betacheck * betacheck
        normr = np.sqrt(delta + (betad - taud) ** 2 + betadd * betadd)
        normA2 += beta * beta
        normA = np.sqrt(normA2)
        normA2 += alpha * alpha
        normAr, normx = abs(zetabar), np.linalg.norm(x)
        if normAr < tol + tol * normA * normr:
            istop = 2
        if normr < tol + tol * (normA * normx + normb):
            istop = 1
    xr, rr, st = oracle.lsmr(a, b, tol, tol)
    assert abs(itn - st["num_steps"]) <= 1 and istop == st["istop"], (itn, st["num_steps"], istop, st["istop"])
    assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
    np.save(os.path.join(out_dir, f"lsmr{rank}.npy"), x)
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_lsmr_decomposition_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_lsmr_rows_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    x0, x1 = np.load(tmp_path / "lsmr0.npy"), np.load(tmp_path / "lsmr1.npy")
    assert np.array_equal(x0, x1)  # replicated state stays bit-identical: the all-reduce result is the same everywhere


def _gmres_rows_worker(rank, world, port, out_dir):
    """The decomposition gmres_dist.cu uses, restated with numpy + gloo: rank p owns a block of rows of A and the same
    slice of every vector; per Arnoldi step an all-gather of the Krylov vector, ONE fused all-reduce of the restart + 1
    Gram-Schmidt projections together with ||w||^2, one all-reduce for the new norm; max-norms by all-reduce(MAX)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import gen
    from oracle.direct import qr_compute, qr_init
    from lineax_b200._shard import shard_bounds

    n, tol, restart, stagnation_iters = 203, 1e-10, 20, 20
    a, b, _ = gen.easy_problem(5, n, np.float64, spd=False)
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    al, bl = a[lo:hi], b[lo:hi]
    eps = np.finfo(np.float64).eps

    def allsum(vec):
        t = torch.as_tensor(np.ascontiguousarray(np.atleast_1d(vec), dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    def allmax(x):
        t = torch.as_tensor(np.array([x], dtype=np.float64))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather(vl):  # all-gather of a row-sharded vector (gloo wants equal sizes: pad to the largest shard)
        width = max(bounds[r + 1] - bounds[r] for r in range(world))
        mine = torch.zeros(width, dtype=torch.float64)
        mine[: hi - lo] = torch.as_tensor(np.ascontiguousarray(vl))
        parts = [torch.empty(width, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, mine)
        return torch.cat([parts[r][: bounds[r + 1] - bounds[r]] for r in range(world)]).numpy()

    def mv(vl):
        return al @ gather(vl)

    def not_converged(rl, dl, yl):  # gmres.py:130-141 with the default max_norm
        rn = allmax(np.max(np.abs(rl / (tol + tol * np.abs(bl)))) if rl.size else 0.0)
        dn = allmax(np.max(np.abs(dl / (tol + tol * np.abs(yl)))) if dl.size else 0.0)
        return rn > 1 or dn > 1

    def main_gmres(yl, rl):
        r_norm = np.sqrt(allsum(rl @ rl)[0])
        initial_breakdown = r_norm < eps
        basis = np.zeros((hi - lo, restart + 1))
        basis[:, 0] = rl / (np.inf if initial_breakdown else r_norm)
        coeff = np.eye(restart, restart + 1)
        breakdown, k = initial_breakdown, 0
        while k < restart and not breakdown:
            w = mv(basis[:, k])
            red = allsum(np.concatenate([basis.T @ w, [w @ w]]))  # fused: projections + ||w||^2
            proj, step_norm = red[:-1], np.sqrt(red[-1])
            w = w - basis @ proj
            nrm = np.sqrt(allsum(w @ w)[0])
            breakdown = bool(nrm < step_norm * eps)
            basis[:, k + 1] = w / (np.inf if breakdown else nrm)
            proj[k + 1] = nrm
            coeff[k, :] = proj
            k += 1
        beta_vec = np.zeros(restart + 1)
        beta_vec[0] = r_norm
        z = qr_compute(qr_init(np.ascontiguousarray(coeff.T)), beta_vec)  # replicated small solve
        diff = basis[:, :-1] @ z
        return yl + diff, diff, breakdown

    with np.errstate(all="ignore"):
        yl, rl = np.zeros(hi - lo), np.zeros(hi - lo)
        breakdown = deferred = False
        diff = np.full(hi - lo, np.inf)
        r_min, step, stag, ms = np.inf, 0, 0, 10 * n
        while ((not deferred) and stag < stagnation_iters and not_converged(rl, diff, yl) and step < ms) or step == 0:
            if step == 0:
                y_new, diff_new, bd = yl, np.full(hi - lo, np.inf), False
            else:
                y_new, diff_new, bd = main_gmres(yl, rl)
            r_new = bl - mv(y_new)
            r_new_norm = allmax(np.max(np.abs(r_new)) if r_new.size else 0.0)
            stag = 0 if (r_new_norm - r_min) < 0 else stag + 1
            r_min = min(r_new_norm, r_min)
            yl, rl, deferred, breakdown, diff = y_new, r_new, breakdown, bd, diff_new
            step += 1
    x = gather(yl)
    xr, rr, st = oracle.gmres(a, b, tol, tol)
    assert step == st["num_steps"] and rr == 0, (step, st["num_steps"], rr)
    assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
    np.save(os.path.join(out_dir, f"gmres{rank}.npy"), x)
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_gmres_decomposition_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_gmres_rows_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    x0, x1 = np.load(tmp_path / "gmres0.npy"), np.load(tmp_path / "gmres1.npy")
    assert np.array_equal(x0, x1)


def _cg_rows_worker(rank, world, port, out_dir):
    """The decomposition cg_dist.cu uses, restated with numpy + gloo: rank p owns a block of rows of the SPD operator and
    the same slice of every vector; per iteration an all-gather of the search direction, an all-reduce of <Ap, p>, and ONE
    round carrying <r, r> with the two max-norms of the convergence test (cg.py:114-227, stabilise_every = 10)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import gen
    from lineax_b200._shard import shard_bounds

    n, tol, stabilise_every = 157, 1e-10, 10
    a, b, _ = gen.easy_problem(7, n, np.float64, spd=True)
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    al, bl = a[lo:hi], b[lo:hi]
    rcond = 2 * np.finfo(np.float64).eps * n

    def allsum(x):
        t = torch.as_tensor(np.array([x], dtype=np.float64))
        dist.all_reduce(t)
        return float(t.item())

    def allmax2(x, y):
        t = torch.as_tensor(np.array([x, y], dtype=np.float64))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    def gather(vl):
        width = max(bounds[r + 1] - bounds[r] for r in range(world))
        mine = torch.zeros(width, dtype=torch.float64)
        mine[: hi - lo] = torch.as_tensor(np.ascontiguousarray(vl))
        parts = [torch.empty(width, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, mine)
        return torch.cat([parts[r][: bounds[r + 1] - bounds[r]] for r in range(world)]).numpy()

    with np.errstate(all="ignore"):
        yl = np.zeros(hi - lo)
        rl = bl - al @ gather(yl)
        pl = rl.copy()
        gamma = allsum(pl @ rl)
        norm1 = norm2 = np.inf
        step, ms = 0, 10 * n
        while gamma > 0 and step < ms and (norm1 > 1 or norm2 > 1):
            ql = al @ gather(pl)
            ip = allsum(ql @ pl)
            alpha = gamma / ip if abs(ip) > 100 * rcond * abs(gamma) else np.nan
            step += 1
            dl = alpha * pl
            yl = yl + dl
            rl = bl - al @ gather(yl) if step % stabilise_every == 0 else rl - alpha * ql
            gn = allsum(rl @ rl)
            norm1, norm2 = allmax2(np.max(np.abs(rl / (tol + tol * np.abs(bl)))), np.max(np.abs(dl / (tol + tol * np.abs(yl)))))
            beta, gamma = gn / gamma, gn
            pl = rl + beta * pl
    x = gather(yl)
    xr, rr, st = oracle.cg(a, b, tol, tol)
    assert rr == 0 and abs(step - st["num_steps"]) <= 1, (step, st["num_steps"], rr)
    assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
    np.save(os.path.join(out_dir, f"cg{rank}.npy"), x)
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_cg_decomposition_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_cg_rows_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    x0, x1 = np.load(tmp_path / "cg0.npy"), np.load(tmp_path / "cg1.npy")
    assert np.array_equal(x0, x1)


def _tsqr_worker(rank, world, port, out_dir):
    """RowShardedQR (TSQR) and its `lx.linear_solve(RowShardedMatrixLinearOperator, b_local, lx.QR())` entry on
    a world-2 gloo group with the oracle-backed CPU doubles: local QR, all-gather of the R factors, small QR."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lineax_b200 as lx
    from lineax_b200 import distributed as lxd
    from oracle import gen
    from tests import cpu_kernels

    cpu_kernels.install()
    lx.set_default_device("cpu")
    m, n = 203, 17
    a, b, _ = gen.tall_lstsq(9, m, n, np.float64)
    solver = lxd.RowShardedQR(m, n, dtype=torch.float64, device="cpu")
    lo, hi = solver.row_range()
    x = solver.solve(torch.as_tensor(a[lo:hi]), torch.as_tensor(b[lo:hi]))
    xl = np.linalg.lstsq(a, b, rcond=None)[0]
    assert np.allclose(x.numpy(), xl, rtol=1e-10, atol=1e-12)
    op = lxd.RowShardedMatrixLinearOperator(torch.as_tensor(a[lo:hi]), m)
    op._solvers[("qr", m, n, torch.float64)] = solver
    sol = lx.linear_solve(op, torch.as_tensor(b[lo:hi]), lx.QR(), throw=False)
    assert np.allclose(sol.value.numpy(), xl, rtol=1e-10, atol=1e-12) and int(sol.result) == 0
    np.save(os.path.join(out_dir, f"q{rank}.npy"), x.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_tsqr_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_tsqr_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert np.array_equal(np.load(tmp_path / "q0.npy"), np.load(tmp_path / "q1.npy")), "replicated bit for bit"
