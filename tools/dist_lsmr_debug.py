import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.cuda.set_device(0)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29541")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
import oracle
from oracle import gen
from lineax_b200 import _ops
from lineax_b200.distributed import RowShardedLSMR
for (m, n) in ((4099, 256), (4096, 2048), (8192, 4096)):
    a, b, _ = gen.tall_lstsq(1, m, n, np.float64)
    A = torch.as_tensor(a).cuda(); B = torch.as_tensor(b).cuda()
    for ms in (1, 2, 3, None):
        solver = RowShardedLSMR(m, n, 1e-12, 1e-12, max_steps=ms, dtype=torch.float64)
        x, res, steps, st = solver.solve(A, B)
        flags = 0 if ms is None else 4
        x1, r1, s1, st1 = _ops.lsmr(A[None], B[None], None, 1e-12, 1e-12, 1e8, 10 * n if ms is None else ms, flags)
        e = float((x - x1[0]).abs().max() / x1[0].abs().max())
        print(m, n, "max_steps", ms, "steps", int(steps), int(s1[0]), "relerr", f"{e:.2e}",
              "stats dist", [f"{float(st[k]):.6e}" for k in ("norm_r", "norm_Ar", "norm_A", "cond_A", "norm_x")],
              "single", [f"{float(v):.6e}" for v in st1[0, 1:6]])
