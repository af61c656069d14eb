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
