"""Run under torchrun on >= 2 GPUs: row-sharded LSMR (peer-memory fused exchange) vs the oracle and
vs the single-GPU grid kernel.  `python -m torch.distributed.run --nproc-per-node 2 tests/dist_lsmr_check.py`"""
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
    from lineax_b200.distributed import RowShardedLSMR

    ok = True
    for (m, n), dtype, tol in (((4099, 256), np.float32, 1e-6), ((16384, 512), np.float32, 1e-6),
                               ((3000, 128), np.float64, 1e-12)):
        a, b, _ = gen.tall_lstsq(m + n, m, n, dtype)
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        solver = RowShardedLSMR(m, n, tol, tol, dtype=tdt)
        lo, hi = solver.row_range()
        A = torch.as_tensor(a).cuda()
        B = torch.as_tensor(b).cuda()
        for rep in range(2):  # twice: the epoch/flag state must carry over between calls
            x, res, steps, stats = solver.solve(A[lo:hi], B[lo:hi])
        torch.cuda.synchronize()
        xs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(xs, x)
        same = all(torch.equal(xs[0], xi) for xi in xs)  # replicated bit for bit
        x = x.cpu().numpy()
        xr, rr, st = oracle.lsmr(a, b, tol, tol)
        x1, r1, s1, st1 = _ops.lsmr(A[None], B[None], None, tol, tol, 1e8, 10 * n, 0)
        err = np.abs(x - xr).max() / np.abs(xr).max()
        err1 = np.abs(x - x1[0].cpu().numpy()).max() / np.abs(xr).max()
        good = (same and int(res) == rr and abs(int(steps) - st["num_steps"]) <= 2
                and int(stats["istop"]) == st["istop"] and err < (1e-5 if dtype == np.float32 else 1e-11))
        ok &= good
        if rank == 0:
            print(f"{m}x{n} {dtype.__name__}: result {int(res)} (oracle {rr}) steps {int(steps)} (oracle "
                  f"{st['num_steps']}, 1-GPU {int(s1[0])}) istop {int(stats['istop'])} rel err vs oracle {err:.2e} "
                  f"vs 1-GPU {err1:.2e} replicated {same} -> {'OK' if good else 'FAIL'}")
    # warm start and max_steps code path
    a, b, _ = gen.tall_lstsq(7, 2048, 128, np.float64)
    solver = RowShardedLSMR(2048, 128, 1e-12, 1e-12, max_steps=3, dtype=torch.float64)
    lo, hi = solver.row_range()
    y0 = np.full(128, 0.5)
    x, res, steps, stats = solver.solve(torch.as_tensor(a[lo:hi]).cuda(), torch.as_tensor(b[lo:hi]).cuda(),
                                        y0=torch.as_tensor(y0).cuda())
    xr, rr, st = oracle.lsmr(a, b, 1e-12, 1e-12, y0=y0, max_steps=3)
    err = np.abs(x.cpu().numpy() - xr).max() / np.abs(xr).max()
    good = int(res) == rr and int(steps) == st["num_steps"] == 3 and err < 1e-9
    ok &= good
    if rank == 0:
        print(f"y0 + max_steps=3: result {int(res)} (oracle {rr}) steps {int(steps)} err {err:.2e} -> "
              f"{'OK' if good else 'FAIL'}")
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_LSMR_ALL_OK" if int(t.item()) == 1 else "DIST_LSMR_FAILED")
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
