"""Run under torchrun on >= 2 GPUs: row-sharded GMRES (peer-memory fused exchanges) vs the oracle
and vs the single-GPU grid kernel.  `python -m torch.distributed.run --nproc-per-node 2 tests/dist_gmres_check.py`"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    import oracle
    from oracle import gen
    from lineax_b200 import _ops
    from lineax_b200.distributed import RowShardedGMRES

    ok = True
    for n, dtype, tol in ((515, np.float32, 1e-6), (2048, np.float32, 1e-6), (1000, np.float64, 1e-12)):
        a, b, _ = gen.easy_problem(n, n, dtype, spd=False)
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        solver = RowShardedGMRES(n, tol, tol, dtype=tdt)
        lo, hi = solver.row_range()
        A = torch.as_tensor(a).cuda()
        B = torch.as_tensor(b).cuda()
        for rep in range(2):  # twice: the epoch/flag state must carry over between calls
            x_loc, res, steps = solver.solve(A[lo:hi], B[lo:hi])
        torch.cuda.synchronize()
        xs = [torch.empty(solver.bounds[r + 1] - solver.bounds[r], dtype=tdt, device="cuda") for r in range(world)]
        dist.all_gather(xs, x_loc)
        x = torch.cat(xs).cpu().numpy()
        xr, rr, st = oracle.gmres(a, b, tol, tol)
        x1, r1, s1 = _ops.gmres(A[None], B[None], None, None, tol, tol, 10 * n, 20, 20, 0)
        err = np.abs(x - xr).max() / np.abs(xr).max()
        err1 = np.abs(x - x1[0].cpu().numpy()).max() / np.abs(xr).max()
        good = int(res) == rr and abs(int(steps) - st["num_steps"]) <= 2 and err < (1e-5 if dtype == np.float32 else 1e-11)
        ok &= good
        if rank == 0:
            print(f"n={n} {dtype.__name__}: result {int(res)} (oracle {rr}) steps {int(steps)} (oracle {st['num_steps']}, "
                  f"1-GPU {int(s1[0])}) rel err vs oracle {err:.2e} vs 1-GPU {err1:.2e} -> {'OK' if good else 'FAIL'}")
    # failure code path: restart=2 on a Gaussian matrix must report a failure on every rank
    rng = np.random.default_rng(0)
    a = rng.standard_normal((600, 600))
    b = a @ rng.standard_normal(600)
    solver = RowShardedGMRES(600, 1e-10, 1e-10, restart=2, dtype=torch.float64)
    lo, hi = solver.row_range()
    x_loc, res, steps = solver.solve(torch.as_tensor(a[lo:hi]).cuda(), torch.as_tensor(b[lo:hi]).cuda())
    xr, rr, st = oracle.gmres(a, b, 1e-10, 1e-10, restart=2)
    good = int(res) == rr and rr != 0
    ok &= good
    if rank == 0:
        print(f"restart=2 failure code: {int(res)} (oracle {rr}) -> {'OK' if good else 'FAIL'}")
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_GMRES_ALL_OK" if int(t.item()) == 1 else "DIST_GMRES_FAILED")
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
