#!/usr/bin/env python3
"""Where does a multi-GPU bench step spend its time?  (name the 1 -> N limiter.)

Run under torchrun like bench.py:  python -m torch.distributed.run --nproc-per-node N ... tools/step_timeline.py
Same workload as bench.py (16384 rows per GPU x 16384, Poisson fp64, depth 4, 32 sweeps per step).  One step is replayed
with CUDA events around every piece of StripSolver.run -- strip-edge row blocks (phase 1, side stream), the NCCL
batch_isend_irecv of the 2*depth halo rows, the interior row blocks (phase 2, main stream), the stream join, and the
MAX all-reduce + device->host read that ends a step -- and is compared with (a) the same passes without the exchange
and (b) one unsplit launch per pass (what a single GPU runs).  Times are the max over ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
from analog_pde_solver.strips import CudaStripEngine, StripSolver, partition_rows, sweep_groups

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows_per_gpu, cols, depth, sweeps = 16384, 16384, 4, 32
grows = rows_per_gpu * world
b, e = partition_rows(grows, world)[rank]
eng = CudaStripEngine(grows, cols, b, e, dtype=np.float64, has_noise=False, has_source=True, temporal_depth=depth)
eng.plan.fill(kind=1, top_hot=False, noise_level=0.0, seed=1234)
sol = StripSolver(eng, rank, world, depth, dist if world > 1 else None)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def rmax(x):
    t = torch.tensor(x, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed_step(mode):
    """mode: 'overlap' (what ships), 'noexchange' (same launches, halo exchange skipped: wrong ghost rows, timing only),
    'unsplit' (one launch per pass, no exchange: the single-GPU schedule)."""
    marks = []
    first = sol.sweeps_done
    t0 = ev(); t0.record()
    for n in sweep_groups(sweeps, depth):
        if mode == "unsplit":
            a, c = ev(), ev()
            a.record(); eng.run(n, sol.sweeps_done, phase=0); c.record()
            marks.append(("pass", a, c))
        else:
            eng.clear_residuals(sol.sweeps_done, n)
            m0 = ev(); m0.record()
            with eng.comm_scope():
                a, c, d = ev(), ev(), ev()
                a.record(); eng.run(n, sol.sweeps_done, phase=1); c.record()
                if mode == "overlap":
                    sol.exchange_halos()
                d.record()
            f, g = ev(), ev()
            f.record(); eng.run(n, sol.sweeps_done, phase=2); g.record()
            eng.join_comm()
            h = ev(); h.record()
            marks.append(("split", m0, a, c, d, f, g, h))
        sol.sweeps_done += n
    r0 = ev(); r0.record()
    sol.residuals(first, sweeps)
    r1 = ev(); r1.record()
    torch.cuda.synchronize()
    out = {"step": t0.elapsed_time(r1), "residuals": r0.elapsed_time(r1)}
    if mode == "unsplit":
        out["pass"] = float(np.mean([a.elapsed_time(c) for _, a, c in marks]))
    else:
        out["edge"] = float(np.mean([m[2].elapsed_time(m[3]) for m in marks]))
        out["exchange"] = float(np.mean([m[3].elapsed_time(m[4]) for m in marks]))
        out["interior"] = float(np.mean([m[5].elapsed_time(m[6]) for m in marks]))
        out["pass"] = float(np.mean([m[1].elapsed_time(m[7]) for m in marks]))
        out["edge_start_delay"] = float(np.mean([m[1].elapsed_time(m[2]) for m in marks]))
        out["join_tail"] = float(np.mean([m[6].elapsed_time(m[7]) for m in marks]))
    return out


for mode in ("overlap", "noexchange", "unsplit"):
    if world == 1 and mode != "unsplit":
        continue
    for _ in range(3):
        timed_step(mode)
    sync()
    acc = {}
    reps = 10
    for _ in range(reps):
        for k, v in timed_step(mode).items():
            acc[k] = acc.get(k, 0.0) + v / reps
    keys = sorted(acc)
    vals = rmax([acc[k] for k in keys])
    if rank == 0:
        print(f"[N={world}] {mode:10s} " + "  ".join(f"{k}={v:.3f}ms" for k, v in zip(keys, vals)), flush=True)
    sync()
if rank == 0:
    print("pass = one fused pass of 4 sweeps (8 per step); edge / interior = kernel time on its stream; exchange = from the end of the "
          "edge kernel to the end of the NCCL send/recv on the side stream; join_tail = from the end of the interior kernel to the join; "
          "residuals = MAX all-reduce of the step's 32 maxima + device->host read", flush=True)
if world > 1:
    dist.destroy_process_group()
