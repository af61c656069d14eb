#!/usr/bin/env python
"""Benchmark of the lineax solve hot path on B200 (contract: see the task statement).

A "step" is one pass of the hot path over one batch of synthetic systems.  The headline line is
BASELINE.json configs[1]: vmapped LU on 65536 independent 32x32 fp32 systems (`--workload lu32`).
The default run ALSO times the other BASELINE configs and attaches them under `"extras"`:
`cg256` (configs[2]), `gmres32k` (configs[3]), `lsmr262k`, `qr262k`, `tridiag512` (configs[4]),
each with its own `roofline`, `cpu_baseline`, `e2e` and an in-run `parity` block (`--no-extras`
skips them, `--workload X` makes X the headline and runs nothing else).

Multi-GPU (`--gpus N`, launched under torchrun, one rank per GPU): vmapped workloads are batch-sharded
(weak scaling, no data-path collective); `gmres32k` / `lsmr262k` row-shard ONE system over the ranks
(strong scaling, exchanges fused into the persistent kernels over NVLink peer memory); `qr262k`
row-shards ONE system too (TSQR: local blocked QR, all-gather of the R factors, small QR).

`--impl reference` times the CPU implementation of the same path (the oracle port: JAX is not
installable in this image) on the host cores.
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
EXTRAS = ("cg256", "tridiag512", "gmres32k", "lsmr262k", "qr262k")
NOMINAL_FP32_TFLOPS = 2 * 148 * 128 * 1.965e9 / 1e12


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_of(workload):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        return json.load(open(tp)).get(workload)
    return None


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


# ------------------------------------------------------------------------------------------ inputs
def make_inputs(workload, seed):
    """Host inputs of the vmapped small-system workloads (lu32, cg256)."""
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


# --------------------------------------------------------------------------------- CPU baselines
def _timed_loop(fn, budget_s, min_reps=1):
    fn()  # warm-up (page faults, BLAS thread pool)
    t0 = time.perf_counter()
    reps = 0
    while reps < min_reps or time.perf_counter() - t0 < budget_s:
        fn()
        reps += 1
    return (time.perf_counter() - t0) / reps, reps


def cpu_baseline(workload, budget_s=6.0, lu_inputs=None):
    """The oracle (CPU restatement of the reference algorithm) timed on the host cores on a bounded
    sample of the same workload; big single systems run at a reduced size of the SAME generator and
    are scaled by the stated work ratio.  Returns the `cpu_baseline` object of the JSON line."""
    import oracle
    from oracle import clib, gen

    cores = os.cpu_count() or 1
    w = WORKLOADS[workload]
    if workload == "lu32":
        a, b = lu_inputs if lu_inputs is not None else make_inputs("lu32", 0)
        clib.lib()
        dt, reps = _timed_loop(lambda: clib.lu_factor_solve(a, b, cores), budget_s)
        return {"value": a.shape[0] / dt, "unit": "solves/s", "cores": cores, "kind": "port",
                "sample": f"{reps} passes over {a.shape[0]} systems (full batch), C getf2/getrs restatement on a "
                          f"{cores}-thread pool; lineax/JAX itself is not installable here"}
    if workload == "cg256":
        a, b, _ = gen.easy_problem(0, 256, np.float32, spd=True, batch=64)
        dt, reps = _timed_loop(lambda: [oracle.cg(a[i], b[i], 1e-6, 1e-6) for i in range(64)], budget_s)
        return {"value": 64 / dt, "unit": "solves/s", "cores": 1, "kind": "port",
                "sample": f"{reps} passes over 64 systems of the same generator, NumPy restatement of cg.py"}
    if workload == "tridiag512":
        d, l, u, b = gen.tridiagonal_systems(0, 2048, 512, np.float32)
        dt, reps = _timed_loop(lambda: [oracle.tridiagonal_compute(d[i], l[i], u[i], b[i]) for i in range(2048)],
                               budget_s)
        return {"value": 2048 / dt, "unit": "solves/s", "cores": 1, "kind": "port",
                "sample": f"{reps} passes over 2048 systems of the same generator, LAPACK sgtsv per system (SciPy)"}
    if workload == "gmres32k":
        n = 4096
        a, b, _ = gen.easy_problem(0, n, np.float32, spd=False)
        dt, reps = _timed_loop(lambda: oracle.gmres(a, b, 1e-6, 1e-6), budget_s)
        scale = (w["n"] / n) ** 2
        return {"value": 1.0 / (dt * scale), "unit": "solves/s", "cores": cores, "kind": "port",
                "sample": f"{reps} solves at n={n} (same generator, same restart count), {dt * 1e3:.1f} ms each, scaled "
                          f"by the operator-size ratio {scale:.0f}x (GEMV-bound); NumPy/OpenBLAS restatement of gmres.py"}
    if workload == "lsmr262k":
        m, n = 32768, 1024
        a, b, _ = gen.tall_lstsq(0, m, n, np.float32)
        dt, reps = _timed_loop(lambda: oracle.lsmr(a, b, 1e-6, 1e-6), budget_s)
        scale = (w["m"] * w["n"]) / (m * n)
        return {"value": 1.0 / (dt * scale), "unit": "solves/s", "cores": cores, "kind": "port",
                "sample": f"{reps} solves at {m}x{n} (same generator), {dt * 1e3:.1f} ms each, scaled by the operator-size "
                          f"ratio {scale:.0f}x (GEMV-bound); NumPy/OpenBLAS restatement of lsmr.py"}
    if workload == "qr262k":
        m, n = 32768, 1024
        a, b, _ = gen.tall_lstsq(0, m, n, np.float32)
        dt, reps = _timed_loop(lambda: oracle.qr_compute(oracle.qr_init(a), b), budget_s)
        fl = lambda mm, nn: 2.0 * mm * nn * nn - 2.0 / 3.0 * nn ** 3 + 4.0 * mm * nn
        scale = fl(w["m"], w["n"]) / fl(m, n)
        return {"value": 1.0 / (dt * scale), "unit": "solves/s", "cores": cores, "kind": "port",
                "sample": f"{reps} solves at {m}x{n} (same generator), {dt * 1e3:.1f} ms each, scaled by the flop ratio "
                          f"{scale:.0f}x; LAPACK sgeqrf + sormqr + strtrs (SciPy/OpenBLAS), what jaxlib's CPU backend calls"}
    raise ValueError(workload)


def native_config(workload, gpus):
    """The `config` object of the native arm (the reference arm reports the very same one)."""
    w = WORKLOADS[workload]
    if workload == "lu32":
        return {"workload": "lu32", "batch_per_gpu": w["batch"], "n": w["n"],
                "parallelism": f"batch-sharded x{gpus}, no collective",
                "l2": "inputs (277 MB per step) larger than the 126 MB L2"}
    return None


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path = the oracle port (JAX is not
    installable), all host threads, W warm-up + exactly K timed passes for lu32."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    if args.workload == "lu32":
        from oracle import clib

        clib.lib()
        cores = os.cpu_count() or 1
        a, b = make_inputs("lu32", 0)
        for _ in range(max(1, args.warmup)):
            clib.lu_factor_solve(a, b, cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            clib.lu_factor_solve(a, b, cores)
        dt = time.perf_counter() - t0
        value = a.shape[0] * args.steps / dt
        cb = {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
              "sample": f"{args.steps} passes over the full batch of {a.shape[0]} systems, C getf2/getrs restatement "
                        f"(oracle/getf2.c) on a {cores}-thread pool; lineax/JAX itself is not installable here"}
        cfg = native_config("lu32", args.gpus)
    else:
        cb = cpu_baseline(args.workload, budget_s=max(2.0, min(20.0, 1.5 * args.steps)))
        value = cb["value"]
        cfg = {"workload": args.workload, "n": w["n"]}
    line = {
        "impl": "reference", "metric": f"batched solves/sec ({w['desc']})", "value": value,
        "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * w["batch"] / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- GPU side
class Ctx:
    def __init__(self, rank, local_rank, world):
        self.rank, self.local_rank, self.world = rank, local_rank, world

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        import torch
        import torch.distributed as dist

        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def time_steps(ctx, step, steps, warmup, soak_s=0.5):
    """W untimed warm-up steps, a short untimed soak so that the clock samples are taken under load,
    then exactly `steps` steps bracketed by barrier + synchronize and per-step CUDA events on the
    launching stream.  Returns (total ms max over ranks, per-launch ms list, launches, clocks)."""
    import torch

    from lineax_b200 import _native as nat

    for _ in range(warmup):
        step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    n_soak = 0
    while time.perf_counter() - t0 < soak_s and n_soak < 200:
        step()
        torch.cuda.synchronize()
        n_soak += 1
    l0 = nat.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ctx.barrier()
    ev[0].record()
    for i in range(steps):
        step()
        ev[i + 1].record()
    ctx.barrier()
    launches = nat.launch_count() - l0
    clocks = sampler.stop() if ctx.rank == 0 else None
    total = ctx.max_over_ranks(ev[0].elapsed_time(ev[-1]))
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return total, per, launches, clocks


def hbm_roofline(alg_bytes, avg_ms, kernel, workload, launches_per_step=1):
    peak, src = measured_peaks()
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic_of(workload), "kernel": kernel, "peak_source": src,
            "algorithmic_bytes_per_launch": int(alg_bytes), "avg_launch_ms": avg_ms,
            "launches_per_step": launches_per_step}


_FP32_PEAK = None


def measured_fp32_peak():
    """FP32 FMA throughput of this GPU, measured with the library's probe kernel (scalar FFMA chains and
    packed fma.rn.f32x2 chains; the larger one is the roofline denominator)."""
    global _FP32_PEAK
    if _FP32_PEAK is not None:
        return _FP32_PEAK
    import torch

    from lineax_b200 import _native as nat

    out = torch.zeros(4, device="cuda")
    fl = ctypes.c_double(0.0)
    res = {}
    stream = torch.cuda.current_stream().cuda_stream
    for packed in (0, 1):
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nat.check(nat.fn("lxb_fp32_fma_probe")(out.data_ptr(), 2000, packed, ctypes.byref(fl), stream), "probe")
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        res["ffma2" if packed else "ffma"] = best
    _FP32_PEAK = {"tflops": max(res.values()), "scalar_ffma_tflops": res["ffma"], "packed_ffma2_tflops": res["ffma2"],
                  "nominal_tflops": NOMINAL_FP32_TFLOPS,
                  "how": "lxb_fp32_fma_probe: 8 independent FMA chains per thread, 64 warps per SM, CUDA events, best of 3"}
    return _FP32_PEAK


# --- vmapped LU ---------------------------------------------------------------------------------
def bench_lu32(ctx, args, with_cpu):
    import torch

    from lineax_b200 import _native as nat
    from oracle import clib

    w = WORKLOADS["lu32"]
    batch, n = w["batch"], w["n"]
    a_h, b_h = make_inputs("lu32", ctx.rank)  # each rank: its own batch (weak scaling)
    A, B = torch.as_tensor(a_h).cuda(), torch.as_tensor(b_h).cuda()
    X = torch.empty_like(B)
    stream = torch.cuda.current_stream().cuda_stream
    fn = nat.fn("lxb_lu_factor_solve_f32")
    argv = (A.data_ptr(), n * n, B.data_ptr(), n, X.data_ptr(), None, None, batch, n, stream)

    def step():
        nat.check(fn(*argv), "lxb_lu_factor_solve_f32")

    total, per, launches, clocks = time_steps(ctx, step, args.steps, args.warmup)
    value = ctx.world * batch * args.steps / (total * 1e-3)
    alg_bytes = batch * (n * n + 2 * n) * 4  # 4352 B per solve, fused, state not written
    roof = hbm_roofline(alg_bytes, float(np.mean(per)), "lu32_tma_kernel (TMA-staged warp-per-system LU, FFMA2)", "lu32")
    # parity: bit-exact against the C oracle on a sample of this very batch
    ns = 2048
    x_ref, _, piv_ref = clib.lu_factor_solve(a_h[:ns], b_h[:ns])
    x_gpu = X[:ns].cpu().numpy()
    LUs = torch.empty(ns, n, n, device="cuda")
    PIV = torch.empty(ns, n, dtype=torch.int32, device="cuda")
    X2 = torch.empty(ns, n, device="cuda")
    nat.check(fn(A.data_ptr(), n * n, B.data_ptr(), n, X2.data_ptr(), LUs.data_ptr(), PIV.data_ptr(), ns, n, stream), "lu")
    parity = {"sample": f"first {ns} systems of the timed batch vs oracle/getf2.c",
              "x_bit_exact": bool(np.array_equal(x_gpu, x_ref)),
              "pivots_bit_exact": bool(np.array_equal(PIV.cpu().numpy(), piv_ref)),
              "max_rel_err": float(np.max(np.abs(x_gpu - x_ref)) / np.max(np.abs(x_ref)))}
    # e2e: host buffers through the C ABI, H2D + D2H inside the timed region
    a_pin, b_pin = torch.as_tensor(a_h).pin_memory(), torch.as_tensor(b_h).pin_memory()
    x_pin = torch.empty(batch, n, dtype=torch.float32).pin_memory()
    nbytes = nat.fn("lxb_host_scratch_bytes")(batch, n, 4)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    hfn = nat.fn("lxb_lu_factor_solve_f32_host")

    def e2e_step():
        nat.check(hfn(a_pin.data_ptr(), b_pin.data_ptr(), x_pin.data_ptr(), batch, n,
                      scratch.data_ptr(), nbytes, stream), "lxb_lu_factor_solve_f32_host")

    for _ in range(2):
        e2e_step()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": ctx.world * batch * e2e_steps / (ms * 1e-3), "unit": "solves/s",
           "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes), "d2h_bytes_per_step": int(batch * n * 4),
           "steps": e2e_steps, "api": "lxb_lu_factor_solve_f32_host (pinned host buffers)"}
    parity["e2e_x_bit_exact"] = bool(np.array_equal(x_pin.numpy()[:ns], x_ref))
    cb = cpu_baseline("lu32", lu_inputs=(a_h, b_h)) if with_cpu and ctx.rank == 0 and ctx.world == 1 else None
    return {
        "metric": f"batched solves/sec ({w['desc']})", "value": value, "unit": "solves/s", "n_gpus": ctx.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": native_config("lu32", ctx.world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb,
        "parity": parity,
    }


# --- vmapped CG ---------------------------------------------------------------------------------
def bench_cg256(ctx, args, with_cpu):
    import torch

    import lineax_b200 as lx
    import oracle
    from lineax_b200 import _native as nat

    w = WORKLOADS["cg256"]
    batch, n = w["batch"], w["n"]
    a_h, b_h = make_inputs("cg256", ctx.rank)
    A, B = torch.as_tensor(a_h).cuda(), torch.as_tensor(b_h).cuda()
    X = torch.empty_like(B)
    RES = torch.empty(batch, dtype=torch.int32, device="cuda")
    STEPS = torch.empty(batch, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    fn = nat.fn("lxb_cg_f32")
    argv = (A.data_ptr(), n * n, B.data_ptr(), n, None, 0, X.data_ptr(), RES.data_ptr(), STEPS.data_ptr(),
            batch, n, 1e-6, 1e-6, 10 * n, 10, 0, None, 0, stream)

    def step():
        nat.check(fn(*argv), "lxb_cg_f32")

    total, per, launches, clocks = time_steps(ctx, step, args.steps, args.warmup)
    value = ctx.world * batch * args.steps / (total * 1e-3)
    k = STEPS.cpu().numpy().astype(np.int64)
    n_mv = 1 + k + k // 10  # operator applications of the reference algorithm
    alg_bytes = int(n_mv.sum()) * n * n * 4
    roof = hbm_roofline(alg_bytes, float(np.mean(per)),
                        "cg_resident_kernel (operator resident in registers + shared memory)", "cg256")
    roof["read_once_bytes"] = int(batch * n * n * 4)
    roof["read_once_frac"] = batch * n * n * 4 / (float(np.mean(per)) * 1e-3) / 1e9 / roof["peak"]
    ns = 16
    xg, kg, rg = X[:ns].cpu().numpy(), k[:ns], RES[:ns].cpu().numpy()
    errs, dsteps, same_res = [], [], True
    for i in range(ns):
        xr, rr, st = oracle.cg(a_h[i], b_h[i], 1e-6, 1e-6)
        errs.append(float(np.abs(xg[i] - xr).max() / np.abs(xr).max()))
        dsteps.append(int(kg[i]) - int(st["num_steps"]))
        same_res &= int(rg[i]) == int(rr)
    parity = {"sample": f"first {ns} systems vs oracle.cg (NumPy restatement of cg.py)", "max_rel_err": max(errs),
              "num_steps_diff_max": int(max(abs(d) for d in dsteps)), "results_equal": bool(same_res),
              "num_steps_gpu": int(kg[0])}
    a_pin, b_pin = torch.as_tensor(a_h).pin_memory(), torch.as_tensor(b_h).pin_memory()
    cg = lx.CG(rtol=1e-6, atol=1e-6)
    solve = torch.func.vmap(lambda m, v: lx.linear_solve(
        lx.MatrixLinearOperator(m, lx.positive_semidefinite_tag), v, cg, throw=False).value)

    def e2e_step():
        return solve(a_pin.cuda(non_blocking=True), b_pin.cuda(non_blocking=True)).cpu()

    e2e_step()
    ctx.barrier()
    e2e_steps = 3
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": ctx.world * batch * e2e_steps / dt, "unit": "solves/s",
           "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes), "d2h_bytes_per_step": int(batch * n * 4),
           "steps": e2e_steps, "api": "torch.func.vmap(lineax_b200.linear_solve(..., CG))"}
    cb = cpu_baseline("cg256") if with_cpu and ctx.rank == 0 and ctx.world == 1 else None
    return {
        "metric": f"batched solves/sec ({w['desc']})", "value": value, "unit": "solves/s", "n_gpus": ctx.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cg256", "batch_per_gpu": batch, "n": n,
                   "parallelism": f"batch-sharded x{ctx.world}, no collective",
                   "l2": "inputs (%.0f MB per step) larger than the 126 MB L2" % ((a_h.nbytes + b_h.nbytes) / 1e6)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb,
        "parity": parity,
    }


# --- vmapped tridiagonal --------------------------------------------------------------------------
def bench_tridiag512(ctx, args, with_cpu):
    import torch

    import oracle
    from lineax_b200 import _ops

    w = WORKLOADS["tridiag512"]
    B_, n = w["batch"], w["n"]
    g = torch.Generator(device="cuda").manual_seed(ctx.rank)
    # strictly diagonally dominant, d = 4 + |N|, |l| + |u| < |d|
    d_ = 4.0 + torch.randn(B_, n, generator=g, device="cuda").abs()
    l_ = torch.randn(B_, n - 1, generator=g, device="cuda").clamp(-1.9, 1.9)
    u_ = torch.randn(B_, n - 1, generator=g, device="cuda").clamp(-1.9, 1.9)
    b = torch.randn(B_, n, generator=g, device="cuda")
    out = {}

    def step():
        out["x"] = _ops.tridiagonal_solve(d_, l_, u_, b)

    steps = max(3, min(args.steps, 10))
    total, per, launches, clocks = time_steps(ctx, step, steps, args.warmup, soak_s=0.3)
    x = out["x"]
    value = ctx.world * B_ * steps / (total * 1e-3)
    roof = hbm_roofline(5 * B_ * n * 4, float(np.mean(per)), "tridiagonal_kernel<float>", "tridiag512",
                        launches_per_step=launches // steps)
    ns = 64
    idx = torch.arange(0, B_, B_ // ns, device="cuda")[:ns]
    dh, lh, uh, bh, xh = (t[idx].cpu().numpy() for t in (d_, l_, u_, b, x))
    errs = []
    for i in range(ns):
        xr = oracle.tridiagonal_compute(dh[i], lh[i], uh[i], bh[i])
        errs.append(float(np.abs(xh[i] - xr).max() / np.abs(xr).max()))
    rr = d_ * x - b
    rr[:, :-1] += u_ * x[:, 1:]
    rr[:, 1:] += l_ * x[:, :-1]
    parity = {"sample": f"{ns} systems spread over the batch vs LAPACK sgtsv (oracle.tridiagonal_compute)",
              "max_rel_err": max(errs), "max_abs_residual_full_batch": float(rr.abs().max())}
    del rr
    pins = [t.cpu().pin_memory() for t in (d_, l_, u_, b)]
    torch.cuda.synchronize()
    ctx.barrier()
    e2e_steps = 2
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dd, ll, uu, bb = (p.cuda(non_blocking=True) for p in pins)
        xo = _ops.tridiagonal_solve(dd, ll, uu, bb).cpu()
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": ctx.world * B_ * e2e_steps / dt, "unit": "solves/s",
           "h2d_bytes_per_step": int(sum(p.numel() for p in pins) * 4), "d2h_bytes_per_step": int(B_ * n * 4),
           "steps": e2e_steps, "api": "lineax_b200._ops.tridiagonal_solve (host pinned -> device -> host)"}
    cb = cpu_baseline("tridiag512") if with_cpu and ctx.rank == 0 and ctx.world == 1 else None
    return {
        "metric": f"batched solves/sec ({w['desc']})", "value": value, "unit": "solves/s", "n_gpus": ctx.world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "tridiag512", "batch_per_gpu": B_, "n": n,
                   "parallelism": f"batch-sharded x{ctx.world}, no collective",
                   "l2": "10.7 GB of operands per step, far larger than the 126 MB L2"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb,
        "parity": parity,
    }


# --- single large systems -------------------------------------------------------------------------
def _reduced_parity(kind):
    """In-run parity of the large-system kernels against the oracle at a size the CPU finishes in a
    second or two, on the SAME generator (the full-size parity runs live in tests/test_fullsize_gpu.py)."""
    import torch

    import oracle
    from lineax_b200 import _ops
    from oracle import gen

    if kind == "gmres":
        n = 2048
        a, b, _ = gen.easy_problem(n + 2, n, np.float32, spd=False)
        x, r, s = _ops.gmres(torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda(), None, None, 1e-6, 1e-6, 10 * n, 20, 20, 0)
        xr, rr, st = oracle.gmres(a, b, 1e-6, 1e-6)
        return {"sample": f"n={n}, same generator, vs oracle.gmres", "result_equal": int(r) == int(rr),
                "max_rel_err": float(np.abs(x.cpu().numpy() - xr).max() / np.abs(xr).max()),
                "num_steps": [int(s), int(st["num_steps"])]}
    m, n = 16384, 256
    a, b, _ = gen.tall_lstsq(3, m, n, np.float32)
    A, Bv = torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda()
    if kind == "lsmr":
        x, r, s, _ = _ops.lsmr(A, Bv, None, 1e-6, 1e-6, 1e8, 10 * n, 0)
        xr, rr, st = oracle.lsmr(a, b, 1e-6, 1e-6)
        return {"sample": f"{m}x{n}, same generator, vs oracle.lsmr", "result_equal": int(r) == int(rr),
                "max_rel_err": float(np.abs(x.cpu().numpy() - xr).max() / np.abs(xr).max()),
                "num_steps": [int(s), int(st["num_steps"])]}
    aq, taus = _ops.qr_factor(A)
    x = _ops.qr_solve(aq, taus, Bv, False)
    xr = oracle.qr_compute(oracle.qr_init(a), b)
    return {"sample": f"{m}x{n}, same generator, vs LAPACK sgeqrf+sormqr+strtrs",
            "max_rel_err": float(np.abs(x.cpu().numpy() - xr).max() / np.abs(xr).max())}


def bench_large(ctx, args, workload, with_cpu):
    import torch
    import torch.distributed as dist

    from lineax_b200 import _ops

    w = WORKLOADS[workload]
    n, m = w["n"], w.get("m", w["n"])
    rank, world = ctx.rank, ctx.world
    g = torch.Generator(device="cuda").manual_seed(1000 + rank)
    force = os.environ.get("LXB_FORCE_SHARDED") == "1"  # experiment: dist kernels on a 1-rank group
    sharded = world > 1 or force
    if force and world == 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", ctx.local_rank))
    gx = torch.Generator(device="cuda").manual_seed(12345)  # the true solution is the same on every rank
    extra = {}
    if workload == "gmres32k":
        xt_full = torch.randn(n, generator=gx, device="cuda", dtype=torch.float32)
        if sharded:
            from lineax_b200.distributed import RowShardedGMRES

            solver = RowShardedGMRES(n, 1e-6, 1e-6, restart=20, dtype=torch.float32)
            lo, hi = solver.row_range()
        else:
            lo, hi = 0, n
        # reference's easy generator (benchmarks/solver_speeds.py:146-152): N(0,1)/n + 2I
        A = torch.randn(hi - lo, n, generator=g, device="cuda", dtype=torch.float32) / n
        A[torch.arange(hi - lo, device="cuda"), torch.arange(lo, hi, device="cuda")] += 2.0
        b = _ops.matvec(A, xt_full, False)
        xt = xt_full[lo:hi]
        if sharded:
            def solve(A_=A, b_=b):
                x, r, k = solver.solve(A_, b_)
                return x, r.reshape(1), k.reshape(1)
            kernel = "gmres_dist_kernel<float> (row-sharded, exchanges fused over NVLink peer memory)"
        else:
            def solve(A_=A, b_=b):
                return _ops.gmres(A_, b_, None, None, 1e-6, 1e-6, 10 * n, 20, 20, 0)
            kernel = "gmres_grid_kernel<float>"
        n_mv = lambda k: 1 + 21 * (k - 1)
        rows_local = hi - lo
    elif workload == "lsmr262k":
        xt = torch.randn(n, generator=gx, device="cuda", dtype=torch.float32)
        if sharded:
            from lineax_b200.distributed import RowShardedLSMR

            solver = RowShardedLSMR(m, n, 1e-6, 1e-6, dtype=torch.float32)
            lo, hi = solver.row_range()
        else:
            lo, hi = 0, m
        A = torch.randn(hi - lo, n, generator=g, device="cuda", dtype=torch.float32) / (m ** 0.5)
        b = _ops.matvec(A, xt, False) + 0.1 * torch.randn(hi - lo, generator=g, device="cuda", dtype=torch.float32)
        if sharded:
            def solve(A_=A, b_=b):
                x, r, k, _ = solver.solve(A_, b_)
                return x, r.reshape(1), k.reshape(1)
            kernel = "lsmr_dist_kernel<float> (row-sharded, exchange fused over NVLink peer memory)"
        else:
            def solve(A_=A, b_=b):
                return _ops.lsmr(A_, b_, None, 1e-6, 1e-6, 1e8, 10 * n, 0)[:3]
            kernel = "lsmr_grid_kernel<float> (fused Golub-Kahan pass)"
        n_mv = lambda k: 2 + 2 * k
        rows_local = hi - lo
    else:  # qr262k
        xt = torch.randn(n, generator=gx, device="cuda", dtype=torch.float32)
        from lineax_b200 import distributed as lxd

        sharded = sharded and hasattr(lxd, "RowShardedQR")
        if sharded:
            solver = lxd.RowShardedQR(m, n, dtype=torch.float32)
            lo, hi = solver.row_range()
        else:
            lo, hi = 0, m
        A = torch.randn(hi - lo, n, generator=g, device="cuda", dtype=torch.float32) / (m ** 0.5)
        b = _ops.matvec(A, xt, False) + 0.1 * torch.randn(hi - lo, generator=g, device="cuda", dtype=torch.float32)
        z = torch.zeros(1, dtype=torch.int32, device="cuda")
        if sharded:
            def solve(A_=A, b_=b):
                return solver.solve(A_, b_), z, z + 1
            kernel = "TSQR: qr_large kernels per rank + all-gather of R + small QR"
        else:
            def solve(A_=A, b_=b):
                aq, taus = _ops.qr_factor(A_)
                return _ops.qr_solve(aq, taus, b_, False), z, z + 1
            kernel = "qr_panel + qr_wbig_ws + qr_update128_ws (two-level blocked Householder, warp-specialised tcgen05 3xTF32 W and trailing update)"
        n_mv = lambda k: 0
        rows_local = hi - lo

    out = {}

    def step():
        out["o"] = solve()

    steps = max(3, min(args.steps, 5 if workload == "qr262k" else 10))
    total, per, launches, clocks = time_steps(ctx, step, steps, args.warmup, soak_s=0.4)
    x, res, k = out["o"]
    k, res = int(k.item()), int(res.item())
    avg_ms = total / steps
    agg = 1 if sharded else world  # row-sharded: ONE system over all ranks; otherwise independent replicas
    value = agg * steps / (total * 1e-3)
    parity = {}
    if workload == "qr262k":
        flops = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3 + 4.0 * m * n
        fp = measured_fp32_peak()
        per_gpu = flops / (world if sharded else 1)
        roof = {"bound": "fp32_fma", "achieved": per_gpu / (avg_ms * 1e-3) / 1e12, "peak": fp["tflops"],
                "unit": "TFLOP/s", "frac": per_gpu / (avg_ms * 1e-3) / 1e12 / fp["tflops"], "traffic": traffic_of(workload),
                "kernel": kernel, "peak_source": "measured on this GPU: " + fp["how"], "fp32_probe": fp,
                "algorithmic_flops_per_step": flops, "avg_launch_ms": avg_ms, "launches_per_step": launches // steps}
        # optimality of the least-squares solution: A^T (A x - b) = 0
        r = _ops.matvec(A, x, False) - b
        gvec = _ops.matvec(A, r, True)
        if sharded:
            dist.all_reduce(gvec)
        parity["full_size"] = {"normal_eq_residual": float(gvec.abs().max() / (r.abs().max() + 1e-30)),
                               "rel_err_vs_xtrue_noise_floor": float((x - xt).abs().max() / xt.abs().max())}
    else:
        alg_bytes = n_mv(k) * rows_local * n * 4  # per GPU (local rows when row-sharded)
        roof = hbm_roofline(alg_bytes, avg_ms, kernel, workload, launches_per_step=launches // steps)
        roof["operator_applications"] = n_mv(k)
        if workload == "gmres32k":
            r = _ops.matvec(A, x if not sharded else _allgather_vec(x, solver), False) - b
            parity["full_size"] = {"rel_residual": float(r.abs().max() / b.abs().max()),
                                   "rel_err_vs_xtrue": float((x - xt).abs().max() / xt.abs().max()),
                                   "num_steps": k, "result": res}
        else:
            r = _ops.matvec(A, x, False) - b
            gvec = _ops.matvec(A, r, True)
            if sharded:
                dist.all_reduce(gvec)
            parity["full_size"] = {"normal_eq_residual": float(gvec.abs().max() / (r.abs().max() + 1e-30)),
                                   "num_steps": k, "result": res}
    if rank == 0:
        parity["reduced"] = _reduced_parity({"gmres32k": "gmres", "lsmr262k": "lsmr", "qr262k": "qr"}[workload])
    # e2e: host operator -> device -> solve -> host solution, every step
    a_pin, b_pin = A.cpu().pin_memory(), b.cpu().pin_memory()
    del A
    torch.cuda.synchronize()
    ctx.barrier()
    e2e_steps = 2
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        Ad, bd = a_pin.cuda(non_blocking=True), b_pin.cuda(non_blocking=True)
        xo = solve(Ad, bd)[0].cpu()
        del Ad
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": agg * e2e_steps / dt, "unit": "solves/s",
           "h2d_bytes_per_step": int(a_pin.numel() * 4 + b_pin.numel() * 4), "d2h_bytes_per_step": int(xo.numel() * 4),
           "steps": e2e_steps, "api": "lineax_b200 solve from pinned host operands (host -> device -> host)"}
    del a_pin
    cb = cpu_baseline(workload) if with_cpu and rank == 0 and world == 1 else None
    return {
        "metric": f"solves/sec ({w['desc']})", "value": value, "unit": "solves/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": avg_ms, "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "rows": m, "cols": n, "rows_per_gpu": rows_local, "num_steps": k, "result": res,
                   "parallelism": (f"row-sharded x{world}, exchanges over NVLink peer memory" if sharded
                                   else f"replica x{world}"),
                   "l2": "the 4.3 GB operator is far larger than the 126 MB L2"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb,
        "parity": parity,
    }


def _allgather_vec(x_loc, solver):
    import torch
    import torch.distributed as dist

    xs = [torch.empty(solver.bounds[r + 1] - solver.bounds[r], dtype=x_loc.dtype, device=x_loc.device)
          for r in range(solver.world)]
    dist.all_gather(xs, x_loc.contiguous())
    return torch.cat(xs)


def run_workload(ctx, args, name, with_cpu):
    import torch

    if name == "lu32":
        r = bench_lu32(ctx, args, with_cpu)
    elif name == "cg256":
        r = bench_cg256(ctx, args, with_cpu)
    elif name == "tridiag512":
        r = bench_tridiag512(ctx, args, with_cpu)
    else:
        r = bench_large(ctx, args, name, with_cpu)
    torch.cuda.synchronize()
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="run ONLY this workload as the headline line (default: lu32 headline + all extras)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    single = args.workload is not None
    args.workload = args.workload or "lu32"
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Ctx(rank, local_rank, world)
    with_cpu = not args.no_cpu_baseline
    line = run_workload(ctx, args, args.workload, with_cpu)
    if not single and not args.no_extras:
        extras = {}
        for name in EXTRAS:
            t0 = time.perf_counter()
            try:
                extras[name] = run_workload(ctx, args, name, with_cpu)
            except Exception as e:  # an extra must never take the headline line down with it
                extras[name] = {"error": f"{type(e).__name__}: {e}"}
            if isinstance(extras[name], dict):
                extras[name]["wall_s"] = round(time.perf_counter() - t0, 1)
        line["extras"] = extras
    if rank == 0:
        print(json.dumps(line))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
