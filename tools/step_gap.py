#!/usr/bin/env python3
"""How much of a bench step is host round trip?  32-sweep steps of the headline workload timed (a) with the blocking
residual read after every step (what bench.py did in round 1), (b) with one read at the end, (c) 320 sweeps in one call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200"))
import numpy as np, torch
from analog_pde_solver import _backend
n, depth, steps = 16384, 3, 10
p = _backend.Plan(n, n, dtype=np.float64, has_noise=False, has_source=True, temporal_depth=depth)
p.fill(kind=1, top_hot=False, seed=7)
def timed(fn):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
base = [0]
def per_step_read():
    for _ in range(steps):
        p.run(32, base[0]); p.residuals(base[0], 32); base[0] += 32
def read_at_end():
    b0 = base[0]
    for _ in range(steps):
        p.run(32, base[0]); base[0] += 32
    p.residuals(b0, 32 * steps)
def one_call():
    p.run(32 * steps, base[0]); p.residuals(base[0], 32 * steps); base[0] += 32 * steps
lups = (n - 2) ** 2 * 32 * steps
for name, fn in (("blocking read per step", per_step_read), ("one read at the end", read_at_end), ("one call, one read", one_call)):
    ms = timed(fn)
    print(f"{name:28s}: {ms / steps:7.3f} ms/step  {lups / ms / 1e6:7.1f} GLUP/s   launches so far {p.launches}")
