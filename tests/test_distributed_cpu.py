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
