#!/usr/bin/env python3
"""bench.py -- GLUP/s (lattice updates per second) of the relaxation hot path on B200.

Workload (BASELINE.json configs[3]): 16384 x 16384 Poisson, red-black Gauss-Seidel, fp64 (or
--dtype f32), temporally blocked streaming kernel.  A *step* = `--sweeps` (default 32 = one
convergence-check interval) full red-black sweeps over the whole grid; inputs are generated on
the device and stay resident in HBM (2 GiB per array, far larger than the 126 MB L2).
N > 1 (torchrun, one rank per GPU): weak scaling -- every rank owns a 16384-row strip of a
(16384*N) x 16384 grid, halo rows move over NCCL once per fused pass, residual maxima are
all-reduced at the end of every step.

One JSON line on rank 0; keys per the driver contract (+ roofline, cpu_baseline, e2e, clocks).

--impl reference : the reference algorithm (lexicographic Gauss-Seidel, solver.py:344-368
upstream) timed on the host cores through the pinned C port under oracle/ (the reference itself
is pure Python and is not on the GPU box), same metric and config, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_LUP = {  # algorithmic bytes per lattice update, one sweep, perfect neighbour reuse
    ("f64", False, False): 16, ("f64", False, True): 24, ("f64", True, False): 32, ("f64", True, True): 40,
    ("f32", False, False): 8, ("f32", False, True): 12, ("f32", True, False): 16, ("f32", True, True): 20,
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--rows", type=int, default=16384, help="rows per GPU")
    ap.add_argument("--cols", type=int, default=16384)
    ap.add_argument("--sweeps", type=int, default=32, help="sweeps per step")
    ap.add_argument("--depth", type=int, default=int(os.environ.get("APDE_TEMPORAL_DEPTH", "0")), help="fused sweeps per HBM pass (0 = measured best: 3, or 2 with mismatch)")
    ap.add_argument("--noise", type=float, default=0.0, help="link mismatch sigma (0 = Poisson, no mismatch)")
    ap.add_argument("--no-source", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-sweeps", type=int, default=256,
                    help="sweeps per end-to-end call (SURVEY 8d config 4: fixed K = 256, check_every = 32)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.depth <= 0:
        args.depth = 2 if args.noise > 0 else 3
    return args


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []   # (arrival time, csv line)
        self.window = [None, None]

    def mark_begin(self):
        self.window[0] = time.time()

    def mark_end(self):
        self.window[1] = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = self.window
        inside = [ln for (ts, ln) in self.lines if t0 is None or (t0 <= ts <= (t1 or ts) + 0.15)]
        if not inside:  # region shorter than one sampling period: take the samples around it
            inside = [ln for (_, ln) in self.lines[-3:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for k, name in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle = test/bench infrastructure; never on the product path)
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(budget_s: float, with_noise: bool, with_source: bool):
    """Reference algorithm (lexicographic GS) on the host cores via the pinned C port, all threads.
    Bounded sample of the bench workload: a 4096 x 4096 tile of the same problem, as many sweeps as
    fit the budget.  Returns (GLUP/s, cores, description)."""
    import oracle
    n = 4096
    threads = oracle.max_threads()
    rng = np.random.default_rng(0)
    g0 = np.zeros((n, n))
    h = 1.0 / (16384 - 1)
    src = None
    if with_source:
        x = np.arange(n) / 16383.0
        src = (h * h) * (-2.0 * np.pi ** 2) * np.outer(np.sin(np.pi * x), np.sin(np.pi * x))
    nh = nv = None
    if with_noise:
        nh = 1.0 + 0.01 * rng.standard_normal((n, n - 1))
        nv = 1.0 + 0.01 * rng.standard_normal((n - 1, n))
    lups = (n - 2) * (n - 2)
    t0 = time.perf_counter()
    oracle.gauss_seidel_mt(g0, nh, nv, src, 2, threads)
    per_sweep = (time.perf_counter() - t0) / 2
    sweeps = int(max(2, min(400, budget_s / max(per_sweep, 1e-6))))
    t0 = time.perf_counter()
    oracle.gauss_seidel_mt(g0, nh, nv, src, sweeps, threads)
    dt = time.perf_counter() - t0
    return lups * sweeps / dt / 1e9, threads, f"{n}x{n} tile of the workload, {sweeps} lexicographic sweeps, {threads} threads (C port of solver.py:344-368, bit-identical to the reference)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    desc = ""
    cores = 1
    per = max(2.0, min(args.cpu_seconds, 120.0 / max(args.steps + args.warmup, 1)))
    for k in range(args.warmup + args.steps):
        v, cores, desc = cpu_reference_sample(per, args.noise > 0, not args.no_source)
        if k >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    out = {
        "impl": "reference", "metric": "GLUP/s", "value": value, "unit": "GLUP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "GLUP/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(args, world):
    kind = "Poisson" if not args.no_source else "Laplace"
    if args.noise > 0:
        kind += f"+{args.noise:g} mismatch"
    return {"workload": f"{args.rows * world}x{args.cols} {kind}, red-black Gauss-Seidel {args.dtype}, "
                        f"{args.sweeps} sweeps/step, temporal depth {args.depth}",
            "rows_per_gpu": args.rows, "cols": args.cols, "sweeps_per_step": args.sweeps,
            "temporal_depth": args.depth, "parallelism": f"row-strips x{world}",
            "l2_policy": "inputs (>= 2 GiB per array) far exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from analog_pde_solver import _backend
    from analog_pde_solver.strips import CudaStripEngine, StripSolver, partition_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dt = np.float64 if args.dtype == "f64" else np.float32
    has_noise, has_src = args.noise > 0, not args.no_source
    grows = args.rows * world
    begin, end = partition_rows(grows, world)[rank]
    eng = CudaStripEngine(grows, args.cols, begin, end, dtype=dt, has_noise=has_noise,
                          has_source=has_src, temporal_depth=args.depth)
    eng.plan.fill(kind=1 if has_src else 0, top_hot=not has_src, noise_level=args.noise, seed=1234)
    solver = StripSolver(eng, rank, world, args.depth, dist if world > 1 else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        first = solver.sweeps_done
        solver.run(args.sweeps)
        return solver.residuals(first, args.sweeps)  # device->host read of the step's result

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.mark_begin()
    l0 = eng.plan.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        res = step()
    ev1.record()
    barrier()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = eng.plan.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    lups_step = (grows - 2) * (args.cols - 2) * args.sweeps
    glups = lups_step * args.steps / (ms * 1e-3) / 1e9

    # ---- single-sweep launches (no temporal blocking): the plain-HBM-roofline reading of the kernel ----
    single = None
    if world == 1:
        n1 = 16
        for _ in range(3):
            eng.run(1, solver.sweeps_done)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n1):
            eng.run(1, solver.sweeps_done)   # one launch = one sweep = one pass over HBM
        s1.record()
        torch.cuda.synchronize()
        single = (grows - 2) * (args.cols - 2) * n1 / (s0.elapsed_time(s1) * 1e-3) / 1e9

    # ---- e2e: HOST buffers through the reference-facing call, copies inside the timed region ---------
    # N = 1: the C-ABI one-shot solve the Python classes make (apde_solve_redblack_f64/f32: upload,
    #        K sweeps with a residual read-back every 32, download), host arrays in pinned memory.
    # N > 1: the strip API (apde_plan_upload_rows / run + halo exchange / apde_plan_download_rows).
    e2e = None
    if not args.no_e2e:
        K = args.e2e_sweeps
        owned = end - begin
        pin = lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
        lo = max(begin - 2 * args.depth, 0)
        hi = min(end + 2 * args.depth, grows)
        host_u = pin((hi - lo, args.cols))
        host_u[:] = 0.0
        host_s = None
        if has_src:
            host_s = pin((hi - lo, args.cols))
            host_s[:] = 1e-9
        host_nh = host_nv = None
        if has_noise:
            host_nh = pin((hi - lo, args.cols - 1))
            host_nh[:] = 1.0
            host_nv = pin((hi - lo - (1 if world == 1 else 0), args.cols))
            host_nv[:] = 1.0
        h2d = host_u.nbytes + (host_s.nbytes if has_src else 0) + (host_nh.nbytes + host_nv.nbytes if has_noise else 0)
        if world == 1:
            mode = "redblack" if args.dtype == "f64" else "redblack32"
            d2h = host_u.nbytes + K * 8

            def e2e_step():
                _backend.solve(mode, host_u, host_nh, host_nv, host_s, 0.0, K, check_every=32,
                               temporal_depth=args.depth)
            api = f"apde_solve_redblack_{args.dtype} (one call = upload + {K} sweeps, residuals read back every 32 + download)"
        else:
            out_host = pin((owned, args.cols))
            d2h = out_host.nbytes + K * 8

            def e2e_step():
                _backend.check(eng.plan._lib.apde_plan_upload_rows(
                    eng.plan._h, lo, hi, _backend._ptr(host_u), _backend._ptr(host_nh), _backend._ptr(host_nv),
                    _backend._ptr(host_s)))
                first = solver.sweeps_done
                solver.run(K)
                solver.residuals(first, K)
                _backend.check(eng.plan._lib.apde_plan_download_rows(eng.plan._h, begin, end, _backend._ptr(out_host)))
            api = f"apde_plan_upload_rows + {K} sweeps with NCCL halo exchange + residual all-reduce + apde_plan_download_rows"

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ts = time.perf_counter()
            e2e_step()
            if os.environ.get("APDE_TRACE"):
                print(f"[bench] e2e call {time.perf_counter() - ts:.3f} s", file=sys.stderr, flush=True)
        barrier()
        wall = time.perf_counter() - t0
        if os.environ.get("APDE_TRACE"):
            print(f"[bench] e2e wall {wall:.3f} s for {args.e2e_steps} calls", file=sys.stderr, flush=True)
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e = {"value": (grows - 2) * (args.cols - 2) * K * args.e2e_steps / float(tw.item()) / 1e9, "unit": "GLUP/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "sweeps_per_call": K,
               "note": api + ", host arrays in pinned memory"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        bpl = BYTES_PER_LUP[(args.dtype, has_noise, has_src)]
        per_gpu_glups = glups / world
        achieved = per_gpu_glups * bpl  # GB/s of algorithmic bytes per GPU
        n_sweep_launches = max(launches, 1)
        traffic = None
        try:  # DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            key = f"{args.dtype}_d{args.depth}_noise{int(has_noise)}_src{int(has_src)}_{args.rows}x{args.cols}"
            traffic = tr.get(key, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_kind": peak_kind, "kernel": "rb_stream_kernel",
                    "algorithmic_bytes_per_lup": bpl,
                    "algorithmic_bytes_per_launch": bpl * (args.rows - 2) * (args.cols - 2) * args.depth,
                    "avg_launch_ms": ms / n_sweep_launches,
                    "single_sweep_launches": None if single is None else {
                        "glups": single, "achieved": single * bpl, "frac": single * bpl / peak,
                        "note": "same kernel, one sweep per launch (no temporal blocking): every launch moves the algorithmic bytes through HBM"},
                    "note": "algorithmic bytes = one read+write of u (+source/links) per sweep; the kernel fuses "
                            f"{args.depth} sweeps per HBM pass, so frac may exceed 1 while DRAM traffic (ncu, profiles/) stays below peak"}
        out = {"metric": "GLUP/s", "value": glups, "unit": "GLUP/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": args.dtype, "data": "synthetic (device-generated sin source, zero ring)",
               "config": workload_config(args, world), "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches),
               "clocks": clocks, "last_residual": float(res[-1])}
        if world == 1 and not args.no_cpu_baseline:
            v, cores, desc = cpu_reference_sample(args.cpu_seconds, has_noise, has_src)
            out["cpu_baseline"] = {"value": v, "unit": "GLUP/s", "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(out), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
