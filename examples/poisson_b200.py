#!/usr/bin/env python3
"""Poisson problem with known solution u = sin(pi x) sin(pi y) on the B200 backend -- the workload of the
reference's examples/poisson_demo.py (20 x 20; BASELINE config 2 also quotes 64 x 64).

    python examples/poisson_b200.py --n 64
    python examples/poisson_b200.py --n 2048 --mode redblack --tol 1e-9 --max-iter 200000
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "analog-pde-solver-sim_b200"))
import numpy as np  # noqa: E402
from analog_pde_solver import AnalogPDESolver, DigitalSolver, EnergyModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20)
ap.add_argument("--mode", default="exact", choices=["exact", "redblack", "redblack32"])
ap.add_argument("--tol", type=float, default=1e-7)
ap.add_argument("--max-iter", type=int, default=10_000)
a = ap.parse_args()
n = a.n
x = np.arange(n) / (n - 1)
exact = np.outer(np.sin(np.pi * x), np.sin(np.pi * x))
f = -2.0 * np.pi ** 2 * exact                      # source f(i, j); solve() scales it by h^2
bc = np.full((n, n), np.nan)
bc[0, :] = bc[-1, :] = 0.0
bc[:, 0] = bc[:, -1] = 0.0
analog, digital = AnalogPDESolver(rows=n, cols=n), DigitalSolver(rows=n, cols=n)
ua = analog.solve(bc, source_fn=f, tol=a.tol, max_iter=a.max_iter, mode=a.mode)
ud = digital.solve(bc, source_fn=f, tol=a.tol, max_iter=a.max_iter, mode=a.mode)
print(f"analog  max error vs exact: {np.max(np.abs(ua - exact)):.4e} ({analog.iterations_run} sweeps)")
print(f"digital max error vs exact: {np.max(np.abs(ud - exact)):.4e} ({digital.iterations_run} sweeps, "
      f"{digital.wall_time_s * 1e3:.1f} ms)")
em = EnergyModel()
ea, ed = analog.energy_joules(em), digital.energy_joules(em)
print(f"analog {ea * 1e12:.3f} pJ, digital {ed * 1e12:.3f} pJ, gain {em.efficiency_ratio(ea, ed):.1f}x")
