#!/usr/bin/env python3
"""Laplace heat plate on the B200 backend -- the workload of the reference's examples/laplace_heat_demo.py
(20 x 20, top edge hot) plus a large-grid variant the reference cannot reach.

    python examples/laplace_heat_b200.py            # 20 x 20, exact (bit-identical) path
    python examples/laplace_heat_b200.py --n 4096 --mode redblack --max-iter 2000
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "analog-pde-solver-sim_b200"))
import numpy as np  # noqa: E402
from analog_pde_solver import AnalogPDESolver, DigitalSolver, EnergyModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20)
ap.add_argument("--mode", default="exact", choices=["exact", "redblack", "redblack32"])
ap.add_argument("--tol", type=float, default=1e-6)
ap.add_argument("--max-iter", type=int, default=20_000)
a = ap.parse_args()
n = a.n

# boundary as an array (NaN = free node): identical to the reference's top_hot callable, but O(1) Python
bc = np.full((n, n), np.nan)
bc[0, :] = 1.0
bc[-1, :] = 0.0
bc[1:, 0] = 0.0
bc[1:, -1] = 0.0

runs = [("analog, ideal", AnalogPDESolver(rows=n, cols=n, noise_level=0.0, rng_seed=42)),
        ("analog, 1 % mismatch", AnalogPDESolver(rows=n, cols=n, noise_level=0.01, rng_seed=0)),
        ("digital", DigitalSolver(rows=n, cols=n))]
sol = {}
for name, s in runs:
    t0 = time.perf_counter()
    sol[name] = s.solve(bc, tol=a.tol, max_iter=a.max_iter, mode=a.mode)
    dt = time.perf_counter() - t0
    lups = (n - 2) ** 2 * s.iterations_run
    print(f"{name:22s}: {s.iterations_run:6d} sweeps, final max_change {s.residual_history[-1]:.2e}, "
          f"{dt * 1e3:8.1f} ms wall, {s.kernel_seconds * 1e3:8.2f} ms kernels "
          f"({lups / max(s.kernel_seconds, 1e-12) / 1e9:.2f} GLUP/s)")
print(f"max |ideal - digital| = {np.max(np.abs(sol['analog, ideal'] - sol['digital'])):.2e}")
print(f"max |noisy - digital| = {np.max(np.abs(sol['analog, 1 % mismatch'] - sol['digital'])):.2e}")
em = EnergyModel(resistor_ohms=10_000, supply_voltage=1.0)
ea, ed = runs[0][1].energy_joules(em), runs[2][1].energy_joules(em)
print(f"analog {ea * 1e12:.3f} pJ, digital {ed * 1e12:.3f} pJ, gain {em.efficiency_ratio(ea, ed):.1f}x")
