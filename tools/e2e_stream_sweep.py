#!/usr/bin/env python3
"""End-to-end time of one apde_solve_redblack call (pinned host arrays, 16384^2 Poisson, 256 sweeps) for several band /
row-block shapes of the streamed schedule, and unstreamed for comparison."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200"))
import numpy as np, torch
from analog_pde_solver import _backend
n, K = 16384, 256
mode = sys.argv[1] if len(sys.argv) > 1 else "redblack"
noise = len(sys.argv) > 2 and sys.argv[2] == "noise"
pin = lambda shape: torch.zeros(shape, dtype=torch.float64).pin_memory().numpy()
u, s = pin((n, n)), pin((n, n)); s[:] = 1e-9
nh = nv = None
if noise:
    nh, nv = pin((n, n - 1)), pin((n - 1, n)); nh[:] = 1.0; nv[:] = 1.0
lups = (n - 2) ** 2 * K
def run(tag):
    best = 1e9
    for _ in range(3):
        u[:] = 0.0
        t0 = time.perf_counter(); hist, it, ksec = _backend.solve(mode, u, nh, nv, s, 0.0, K, check_every=32); dt = time.perf_counter() - t0
        best = min(best, dt)
    print(f"{mode}{' +mismatch' if noise else ''} {tag:28s}: {best*1e3:7.1f} ms  {lups/best/1e9:6.1f} GLUP/s  (kernel span {ksec*1e3:.1f} ms) chk {float(u[n//2, n//2]):.6e} {float(hist[-1]):.6e}", flush=True)
os.environ["APDE_STREAMED"] = "0"; run("unstreamed")
os.environ["APDE_STREAMED"] = "1"
for band, rbh in ((2048, 256), (4096, 512), (4096, 256), (2048, 128), (1024, 128), (3072, 384), (1536, 192), (2048, 512)):
    os.environ["APDE_STREAM_BAND"], os.environ["APDE_STREAM_RBH"] = str(band), str(rbh)
    run(f"band {band} rbh {rbh}")
