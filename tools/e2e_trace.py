#!/usr/bin/env python3
"""End-to-end one-shot solve from pinned host memory with the library's own phase timing (APDE_TRACE).
usage: python tools/e2e_trace.py [f64|f32] [K]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200"))
import numpy as np, torch
from analog_pde_solver import _backend
dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = 16384
os.environ["APDE_TRACE"] = "1"
pin = lambda: torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy()
u, s = pin(), pin()
mode = "redblack" if dt == "f64" else "redblack32"
for rep in range(3):
    u[:] = 0.0; s[:] = 1e-9
    t0 = time.perf_counter()
    hist, it, ksec = _backend.solve(mode, u, None, None, s, 0.0, K, check_every=32)
    dtm = time.perf_counter() - t0
    print(f"{dt} K={K}: {dtm*1e3:7.1f} ms  -> {(n-2)**2*K/dtm/1e9:6.1f} GLUP/s e2e   (kernels {ksec*1e3:.1f} ms)", flush=True)
