#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL):
the row-strip run with halo exchange over NVLink equals the single-GPU run bit for bit.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/strips_check.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from analog_pde_solver import _backend
from analog_pde_solver.strips import CudaStripEngine, StripSolver, partition_rows
from helpers import random_problem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok_all = True
for dt, rows, cols, depth, sweeps, noise in [(np.float64, 4096, 3000, 3, 11, 0.01), (np.float32, 4096, 3000, 3, 11, 0.01),
                                             (np.float64, 2048 * world, 2048, 2, 9, 0.0), (np.float32, 1024 * world, 4100, 4, 8, 0.0)]:
    g0, nh, nv, s = random_problem(rows, cols, noise, seed=17)
    begin, end = partition_rows(rows, world)[rank]
    for overlap in (False, True):
        eng = CudaStripEngine(rows, cols, begin, end, dtype=dt, has_noise=noise > 0, has_source=True, temporal_depth=depth)
        eng.plan.upload(g0, nh, nv, s)
        solver = StripSolver(eng, rank, world, depth, dist)
        solver.run(sweeps, overlap=overlap)
        res = solver.residuals(0, sweeps)
        mine = np.zeros((rows, cols))
        eng.plan.download(mine)
        torch.cuda.synchronize()
        with _backend.Plan(rows, cols, dtype=dt, has_noise=noise > 0, has_source=True, temporal_depth=depth) as whole:
            whole.upload(g0, nh, nv, s)
            done = 0
            while done < sweeps:
                n = min(depth, sweeps - done)
                whole.run(n, done)
                done += n
            ref = whole.download()
            ref_hist = whole.residuals(0, sweeps)
        ok = mine[begin:end].tobytes() == ref[begin:end].tobytes() and res.tobytes() == ref_hist.tobytes()
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{np.dtype(dt).name} {rows}x{cols} depth={depth} sweeps={sweeps} noise={noise} overlap={overlap}: "
                  f"{'bit-exact' if t.item() == 1 else 'MISMATCH'} (exchanges={solver.exchanges})", flush=True)
        ok_all = ok_all and t.item() == 1
        eng.close()
dist.destroy_process_group()
if rank == 0:
    print("STRIPS_CHECK", "PASS" if ok_all else "FAIL", flush=True)
sys.exit(0 if ok_all else 1)
