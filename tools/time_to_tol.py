#!/usr/bin/env python3
"""Time-to-tolerance of the relaxation modes on the GPU: Poisson problem with the known
solution sin(pi x) sin(pi y) (examples/poisson_demo.py upstream), zero ring, start from zero.

For each mode: wall time of the one-shot solve() call (upload + sweeps/cycles + download), kernel time, sweeps or
cycles, and the error against the exact PDE solution -- a solve is only 'done' when that error has reached the
discretisation error ~0.82 h^2, which the reference's own stopping rule (max_change < tol per sweep) does not
guarantee on a large grid.  usage: python tools/time_to_tol.py [N ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200"))
import numpy as np
from analog_pde_solver import DigitalSolver, _backend

sizes = [int(x) for x in sys.argv[1:]] or [1025, 4096, 16384]
print(f"{'N':>6} {'mode':<14} {'tol':>8} {'its':>8} {'wall s':>9} {'kernel s':>9} {'max err vs exact':>17} {'0.82 h^2':>10}")
for n in sizes:
    x = np.arange(n) / (n - 1)
    exact = np.outer(np.sin(np.pi * x), np.sin(np.pi * x))
    f = -2.0 * np.pi ** 2 * exact
    h2 = 0.82 / (n - 1) ** 2
    runs = [("multigrid", 1e-9, 60)]
    if n <= 8192:
        runs += [("redblack_sor", 1e-9, 40 * n), ("redblack_sor", 1e-12, 40 * n)]
        # plain Gauss-Seidel for comparison: capped, it needs ~0.5 N^2 sweeps
        runs += [("redblack", 1e-9, min(200_000, 20 * n))]
    for mode, tol, cap in runs:
        d = DigitalSolver(rows=n, cols=n)
        g0 = np.zeros((n, n))
        t0 = time.perf_counter()
        u = d.solve(g0, source_fn=f, tol=tol, max_iter=cap, mode=mode)
        wall = time.perf_counter() - t0
        err = float(np.max(np.abs(u - exact)))
        flag = "" if d.iterations_run < cap else "  (cap reached, not converged)"
        print(f"{n:>6} {mode:<14} {tol:>8.0e} {d.iterations_run:>8} {wall:>9.3f} {d.kernel_seconds:>9.3f} {err:>17.3e} {h2:>10.2e}{flag}", flush=True)
