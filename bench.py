#!/usr/bin/env python
"""Benchmark of the lineax solve hot path on B200 (contract: see the task statement).

A "step" is one pass of the hot path over one batch of synthetic systems.
Default workload = BASELINE.json configs[1]: vmapped LU on 65536 independent 32x32 fp32
systems (`--workload lu32`); `--workload cg256` is configs[2] (4096 x 256^2 SPD fp32 CG).
Multi-GPU (`--gpus N`, launched under torchrun): batch-sharded, no data-path collective,
every rank solves its own full-size batch (weak scaling).

`--impl reference` times the CPU implementation of the same path (the oracle port: JAX is
not installable in this image) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "lu32": dict(batch=65536, n=32, desc="vmapped lx.LU on 65536 independent 32x32 fp32 systems"),
    "cg256": dict(batch=4096, n=256, desc="vmapped lx.CG on 4096 independent 256x256 SPD fp32 systems, rtol=atol=1e-6"),
    "gmres32k": dict(batch=1, n=32768, desc="lx.GMRES restart=20 on a 32768x32768 nonsymmetric fp32 dense system, rtol=atol=1e-6"),
    "lsmr262k": dict(batch=1, n=4096, m=262144, desc="lx.LSMR least squares on a 262144x4096 tall fp32 matrix, rtol=atol=1e-6"),
    "qr262k": dict(batch=1, n=4096, m=262144, desc="lx.QR least squares (geqrf + ormqr + trtrs) on a 262144x4096 tall fp32 matrix"),
    "tridiag512": dict(batch=1 << 20, n=512, desc="vmapped lx.Tridiagonal on 2^20 independent systems of length 512, fp32"),
}
LARGE = ("gmres32k", "lsmr262k", "qr262k", "tridiag512")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if "Active" in v and "Not" not in v:
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def make_inputs(workload, seed):
    from oracle import gen

    w = WORKLOADS[workload]
    if workload == "lu32":
        rng = np.random.default_rng(seed)
        a = rng.standard_normal((w["batch"], w["n"], w["n"]), dtype=np.float32)
        x = rng.standard_normal((w["batch"], w["n"]), dtype=np.float32)
        b = np.einsum("bij,bj->bi", a, x)
        return a, b
    # cg256: the reference's own easy generator (benchmarks/solver_speeds.py:146-152), 256 distinct
    # matrices tiled 16x (generation of 4096 matrices on the host would dominate the run time)
    a, b, _ = gen.easy_problem(seed, w["n"], np.float32, spd=True, batch=256)
    reps = w["batch"] // 256
    return np.tile(a, (reps, 1, 1)), np.tile(b, (reps, 1))


def cpu_port_step(workload, a, b, threads=None):
    """One pass of the oracle (CPU restatement) over a sample; returns systems solved."""
    from oracle import clib
    import oracle

    if workload == "lu32":
        clib.lu_factor_solve(a, b, threads)
        return a.shape[0]
    for i in range(a.shape[0]):
        oracle.cg(a[i], b[i], 1e-6, 1e-6)
    return a.shape[0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import clib

    clib.lib()
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample = w["batch"] if args.workload == "lu32" else 256
    a, b = make_inputs(args.workload, 0)
    a, b = a[:sample], b[:sample]
    for _ in range(args.warmup):
        cpu_port_step(args.workload, a, b)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += cpu_port_step(args.workload, a, b)
    dt = time.perf_counter() - t0
    value = done / dt
    line = {
        "impl": "reference", "metric": f"batched solves/sec ({w['desc']})", "value": value,
        "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_step": sample, "n": w["n"]},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores if args.workload == "lu32" else 1,
                         "kind": "port",
                         "sample": f"{sample} systems per step; oracle port (C getf2 for LU on a thread pool / "
                                   "NumPy CG), JAX unavailable so lineax itself cannot run"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_large(args, w, rank, local_rank, world):
    """Single large systems (BASELINE configs[3], configs[4]).  gmres32k with N > 1: ONE system
    row-sharded over the ranks (strong scaling, exchanges fused into the kernel over NVLink);
    lsmr262k with N > 1: ONE tall system row-sharded the same way; qr262k / tridiag512: replicas."""
    import torch
    import torch.distributed as dist

    from lineax_b200 import _native as nat
    from lineax_b200 import _ops

    n = w["n"]
    m = w.get("m", n)
    g = torch.Generator(device="cuda").manual_seed(rank)
    force = os.environ.get("LXB_FORCE_SHARDED") == "1"  # experiment: dist kernel on a 1-rank group
    sharded = args.workload in ("gmres32k", "lsmr262k") and (world > 1 or force)
    if force and world == 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
    if sharded and args.workload == "lsmr262k":
        # ONE tall system row-partitioned over the ranks; the A^T u all-reduce is fused in the kernel
        from lineax_b200.distributed import RowShardedLSMR

        solver = RowShardedLSMR(m, n, 1e-6, 1e-6, dtype=torch.float32)
        lo, hi = solver.row_range()
        A = torch.randn(hi - lo, n, generator=g, device="cuda", dtype=torch.float32) / (m ** 0.5)
        gx = torch.Generator(device="cuda").manual_seed(12345)
        xt = torch.randn(n, generator=gx, device="cuda", dtype=torch.float32)
        b = _ops.matvec(A, xt, False) + 0.1 * torch.randn(hi - lo, generator=g, device="cuda", dtype=torch.float32)
        m = hi - lo

        def solve():
            x, r, k, _ = solver.solve(A, b)
            return x, r.reshape(1), k.reshape(1)

        n_mv = lambda k: 2 + 2 * k
        kernel_name = "lsmr_dist_kernel<float>"
    elif sharded:
        # ONE system row-partitioned over the ranks (strong scaling); exchanges fused in the kernel
        from lineax_b200.distributed import RowShardedGMRES

        solver = RowShardedGMRES(n, 1e-6, 1e-6, restart=20, dtype=torch.float32)
        lo, hi = solver.row_range()
        A = torch.randn(hi - lo, n, generator=g, device="cuda", dtype=torch.float32) / n
        A[torch.arange(hi - lo, device="cuda"), torch.arange(lo, hi, device="cuda")] += 2.0
        gx = torch.Generator(device="cuda").manual_seed(12345)
        xt_full = torch.randn(n, generator=gx, device="cuda", dtype=torch.float32)
        b = _ops.matvec(A, xt_full, False)
        xt = xt_full[lo:hi]
        m = hi - lo

        def solve():
            x, r, k = solver.solve(A, b)
            return x, r.reshape(1), k.reshape(1)

        n_mv = lambda k: 1 + 21 * (k - 1)
        kernel_name = "gmres_dist_kernel<float>"
    elif args.workload == "gmres32k":
        # reference's easy generator (benchmarks/solver_speeds.py:146-152): N(0,1)/n + 2I
        A = torch.randn(n, n, generator=g, device="cuda", dtype=torch.float32) / n
        A.diagonal().add_(2.0)
        xt = torch.randn(n, generator=g, device="cuda", dtype=torch.float32)
        b = _ops.matvec(A, xt, False)
        solve = lambda: _ops.gmres(A, b, None, None, 1e-6, 1e-6, 10 * n, 20, 20, 0)
        n_mv = lambda k: 1 + 21 * (k - 1)
        kernel_name = "gmres_grid_kernel<float>"
    elif args.workload == "tridiag512":
        # strictly diagonally dominant, d = 4 + |N|, |l| + |u| < |d|
        B_ = w["batch"]
        d_ = 4.0 + torch.randn(B_, n, generator=g, device="cuda").abs()
        l_ = torch.randn(B_, n - 1, generator=g, device="cuda").clamp(-1.9, 1.9)
        u_ = torch.randn(B_, n - 1, generator=g, device="cuda").clamp(-1.9, 1.9)
        b = torch.randn(B_, n, generator=g, device="cuda")
        A = d_
        xt = None
        m = B_

        def solve():
            x = _ops.tridiagonal_solve(d_, l_, u_, b)
            z = torch.zeros(1, dtype=torch.int32, device="cuda")
            return x, z, z + 1

        n_mv = lambda k: 0
        kernel_name = "tridiagonal_kernel<float>"
    elif args.workload == "qr262k":
        A = torch.randn(m, n, generator=g, device="cuda", dtype=torch.float32) / (m ** 0.5)
        xt = torch.randn(n, generator=g, device="cuda", dtype=torch.float32)
        b = _ops.matvec(A, xt, False) + 0.1 * torch.randn(m, generator=g, device="cuda", dtype=torch.float32)

        def solve():
            aq, taus = _ops.qr_factor(A)
            x = _ops.qr_solve(aq, taus, b, False)
            z = torch.zeros(1, dtype=torch.int32, device="cuda")
            return x, z, z + 1

        n_mv = lambda k: 0
        kernel_name = "qr_panel_kernel + qr_wpartial_kernel + qr_update_kernel (blocked Householder)"
    else:
        A = torch.randn(m, n, generator=g, device="cuda", dtype=torch.float32) / (m ** 0.5)
        xt = torch.randn(n, generator=g, device="cuda", dtype=torch.float32)
        b = _ops.matvec(A, xt, False) + 0.1 * torch.randn(m, generator=g, device="cuda", dtype=torch.float32)
        solve = lambda: _ops.lsmr(A, b, None, 1e-6, 1e-6, 1e8, 10 * n, 0)
        n_mv = lambda k: 2 + 2 * k
        kernel_name = "lsmr_grid_kernel<float>"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = solve()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_soak = time.perf_counter()
    n_soak = 0
    soak_s = 0.0 if os.environ.get("LXB_NO_SOAK") == "1" else 0.4
    while time.perf_counter() - t_soak < soak_s and n_soak < 50:  # clock samples under load (untimed)
        out = solve()
        torch.cuda.synchronize()
        n_soak += 1
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = solve()
    e1.record()
    barrier()
    launches = nat.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    k = int(out[2].item())
    res = int(out[1].item())
    if args.workload == "tridiag512":
        x_ = out[0]
        rr = d_ * x_ - b
        rr[:, :-1] += u_ * x_[:, 1:]
        rr[:, 1:] += l_ * x_[:, :-1]
        xerr = float(rr.abs().max())  # max residual
        alg_bytes = 5 * m * n * 4    # 5 n s per solve
    else:
        xerr = float((out[0] - xt).abs().max() / xt.abs().max())
        alg_bytes = n_mv(k) * m * n * 4  # per GPU (m = local rows when row-sharded)
    peak, peak_src = measured_peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:
        traffic = json.load(open(tp)).get(args.workload)
    achieved = alg_bytes / (ms / args.steps * 1e-3) / 1e9
    # e2e: host matrix -> device -> solve -> host solution, every step
    a_pin = A.cpu().pin_memory()
    b_pin = b.cpu().pin_memory()
    if args.workload == "tridiag512":
        l_pin, u_pin = l_.cpu().pin_memory(), u_.cpu().pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps = 2
    for _ in range(e2e_steps):
        Ad, bd = a_pin.cuda(non_blocking=True), b_pin.cuda(non_blocking=True)
        if args.workload == "tridiag512":
            xo = _ops.tridiagonal_solve(Ad, l_pin.cuda(non_blocking=True), u_pin.cuda(non_blocking=True), bd).cpu()
        elif args.workload == "qr262k":
            aq_, t_ = _ops.qr_factor(Ad)
            xo = _ops.qr_solve(aq_, t_, bd, False).cpu()
        elif sharded:
            xo = solver.solve(Ad, bd)[0].cpu()
        elif args.workload == "gmres32k":
            xo = _ops.gmres(Ad, bd, None, None, 1e-6, 1e-6, 10 * n, 20, 20, 0)[0].cpu()
        else:
            xo = _ops.lsmr(Ad, bd, None, 1e-6, 1e-6, 1e8, 10 * n, 0)[0].cpu()
    dt = time.perf_counter() - t0
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": f"solves/sec ({w['desc']})",
        "value": (1 if sharded else world) * w["batch"] * args.steps / (ms * 1e-3), "unit": "solves/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "rows": m, "cols": n, "num_steps": k, "result": res,
                   "rel_err_vs_xtrue": xerr,
                   "parallelism": (f"row-sharded x{world}, exchanges fused over NVLink peer memory" if sharded
                                   else f"replica x{world}"),
                   "l2": "the 4.3 GB operator is far larger than the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": (1 if sharded else world) * w["batch"] * e2e_steps / dt, "unit": "solves/s",
                "h2d_bytes_per_step": int(A.numel() * 4 + b.numel() * 4 + (2 * l_.numel() * 4 if args.workload == "tridiag512" else 0)),
                "d2h_bytes_per_step": int(n * 4 * w["batch"]), "steps": e2e_steps, "api": "lineax_b200._ops (host pinned -> device -> host)"},
        "gpu_launches": int(launches),
        "roofline": ({"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                      "traffic": traffic, "kernel": kernel_name, "peak_source": peak_src,
                      "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": ms / args.steps}
                     if args.workload != "qr262k" else
                     {"bound": "fp32_fma", "achieved": (2.0 * m * n * n - 2.0 / 3.0 * n ** 3 + 4.0 * m * n) / (ms / args.steps * 1e-3) / 1e12,
                      "peak": 2 * 148 * 128 * 1.965e9 / 1e12, "unit": "TFLOP/s",
                      "frac": (2.0 * m * n * n - 2.0 / 3.0 * n ** 3 + 4.0 * m * n) / (ms / args.steps * 1e-3) / (2 * 148 * 128 * 1.965e9),
                      "traffic": None, "kernel": kernel_name,
                      "peak_source": "nominal fp32 FMA peak 148 SMs x 128 lanes x 2 x 1.965 GHz (not in MEASURED_PEAKS.json)",
                      "algorithmic_flops_per_step": 2.0 * m * n * n - 2.0 / 3.0 * n ** 3 + 4.0 * m * n,
                      "avg_launch_ms": ms / args.steps}),
        "cpu_baseline": None,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="lu32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import lineax_b200 as lx
    from lineax_b200 import _native as nat
    from lineax_b200 import _ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = WORKLOADS[args.workload]
    if args.workload in LARGE:
        return run_large(args, w, rank, local_rank, world)
    batch, n = w["batch"], w["n"]
    a_h, b_h = make_inputs(args.workload, rank)  # each rank: its own batch (weak scaling)
    A = torch.as_tensor(a_h).cuda()
    B = torch.as_tensor(b_h).cuda()
    stream = torch.cuda.current_stream().cuda_stream

    if args.workload == "lu32":
        X = torch.empty_like(B)
        fn = nat.fn("lxb_lu_factor_solve_f32")
        argv = (A.data_ptr(), n * n, B.data_ptr(), n, X.data_ptr(), None, None, batch, n, stream)

        def step():
            nat.check(fn(*argv), "lxb_lu_factor_solve_f32")

        alg_bytes = batch * (n * n + 2 * n) * 4  # 4352 B per solve, fused, state not written
        kernel_name = "lu_warp_kernel<float,32,true>"
    else:
        X = torch.empty_like(B)
        RES = torch.empty(batch, dtype=torch.int32, device="cuda")
        STEPS = torch.empty(batch, dtype=torch.int32, device="cuda")
        fn = nat.fn("lxb_cg_f32")
        argv = (A.data_ptr(), n * n, B.data_ptr(), n, None, 0, X.data_ptr(), RES.data_ptr(), STEPS.data_ptr(),
                batch, n, 1e-6, 1e-6, 10 * n, 10, 0, None, 0, stream)

        def step():
            nat.check(fn(*argv), "lxb_cg_f32")

        alg_bytes = None  # depends on the iteration counts, filled in after the run
        kernel_name = "cg_resident_kernel (operator resident in registers + shared memory)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # The K timed steps last only a few ms, too short for nvidia-smi's sampling period: keep the GPU
    # under the SAME load (untimed) for >= 0.6 s first so the clock samples are taken under load.
    t_soak = time.perf_counter()
    while time.perf_counter() - t_soak < 0.6:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    launches0 = nat.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = nat.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * batch * args.steps / (total_ms_max * 1e-3)

    if args.workload == "cg256":
        k = STEPS.cpu().numpy().astype(np.int64)
        n_mv = 1 + k + k // 10  # operator applications of the reference algorithm
        alg_bytes = int(n_mv.sum()) * n * n * 4

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ----
    e2e = None
    if args.workload == "lu32":
        a_pin = torch.as_tensor(a_h).pin_memory()
        b_pin = torch.as_tensor(b_h).pin_memory()
        x_pin = torch.empty(batch, n, dtype=torch.float32).pin_memory()
        nbytes = nat.fn("lxb_host_scratch_bytes")(batch, n, 4)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        hfn = nat.fn("lxb_lu_factor_solve_f32_host")

        def e2e_step():
            nat.check(hfn(a_pin.data_ptr(), b_pin.data_ptr(), x_pin.data_ptr(), batch, n,
                          scratch.data_ptr(), nbytes, stream), "lxb_lu_factor_solve_f32_host")

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            e2e_step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t2 = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * batch * e2e_steps / (float(t2.item()) * 1e-3), "unit": "solves/s",
               "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes), "d2h_bytes_per_step": int(batch * n * 4),
               "steps": e2e_steps, "api": "lxb_lu_factor_solve_f32_host (pinned host buffers)"}
        x_check = x_pin.numpy().copy()
    else:
        a_pin = torch.as_tensor(a_h).pin_memory()
        b_pin = torch.as_tensor(b_h).pin_memory()
        cg = lx.CG(rtol=1e-6, atol=1e-6)
        solve = torch.func.vmap(lambda m, v: lx.linear_solve(
            lx.MatrixLinearOperator(m, lx.positive_semidefinite_tag), v, cg, throw=False).value)

        def e2e_step():
            return solve(a_pin.cuda(non_blocking=True), b_pin.cuda(non_blocking=True)).cpu()

        e2e_step()
        barrier()
        e2e_steps = 3
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            x_check = e2e_step().numpy()
        barrier()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * batch * e2e_steps / float(t2.item()), "unit": "solves/s",
               "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes), "d2h_bytes_per_step": int(batch * n * 4),
               "steps": e2e_steps, "api": "torch.func.vmap(lineax_b200.linear_solve(..., CG))"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    avg_launch_s = float(np.mean(per_launch_ms)) * 1e-3
    achieved = alg_bytes / avg_launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_launch_s * 1e3}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        if args.workload == "lu32":
            sa, sb = a_h, b_h
            used = cores
        else:
            sa, sb = a_h[:64], b_h[:64]
            used = 1
        cpu_port_step(args.workload, sa[:64], sb[:64])
        t0 = time.perf_counter()
        done = 0
        while time.perf_counter() - t0 < 10.0:
            done += cpu_port_step(args.workload, sa, sb)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": done / dt, "unit": "solves/s", "cores": used, "kind": "port",
                        "sample": f"{done} solves in {dt:.1f}s: oracle port of the same workload "
                                  f"({sa.shape[0]} systems per pass); lineax/JAX itself is not installable here"}

    line = {
        "metric": f"batched solves/sec ({w['desc']})", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": batch, "n": n,
                   "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": "inputs (%.0f MB per step) larger than the 126 MB L2" % ((a_h.nbytes + b_h.nbytes) / 1e6)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
