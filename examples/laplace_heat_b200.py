#!/usr/bin/env python3
"""Laplace heat plate on the B200 backend -- the workload of the reference's examples/laplace_heat_demo.py
(20 x 20, top edge hot: iterations, accuracy, energy model, ASCII heat-map, centre-column profile) at any size.

    python examples/laplace_heat_b200.py                                   # 20 x 20, exact (bit-identical) path
    python examples/laplace_heat_b200.py --n 4096 --mode multigrid --tol 1e-9            # ~12 cycles per solve
    python examples/laplace_heat_b200.py --n 4096 --mode redblack_kcl --omega auto --max-iter 40000

With 1 % mismatch the reference's update diverges beyond N ~ 200 (SURVEY finding 2); --mode redblack_kcl and
--mode multigrid use the normalised (Kirchhoff) update, which converges at every size.  Large grids are shown
block-averaged.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "analog-pde-solver-sim_b200"))
import numpy as np  # noqa: E402
from analog_pde_solver import AnalogPDESolver, DigitalSolver, EnergyModel, _backend  # noqa: E402
from analog_pde_solver.report import ascii_heatmap, centre_column_profile  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20)
ap.add_argument("--mode", default="exact", choices=list(_backend.MODES))
ap.add_argument("--tol", type=float, default=1e-6)
ap.add_argument("--max-iter", type=int, default=20_000)
ap.add_argument("--omega", default=None, help="over-relaxation factor for the sor / kcl modes ('auto' = ideal-mesh optimum)")
ap.add_argument("--no-maps", action="store_true")
a = ap.parse_args()
n = a.n
omega = None if a.omega is None else (_backend.optimal_omega(n, n) if a.omega == "auto" else float(a.omega))

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
    mode = a.mode
    kw = {"omega": omega} if (omega is not None and ("sor" in mode or "kcl" in mode)) else {}
    if mode.startswith("redblack_kcl") and "mismatch" in name and not kw:
        kw = {"omega": _backend.optimal_omega(n, n)}
    t0 = time.perf_counter()
    sol[name] = s.solve(bc, tol=a.tol, max_iter=a.max_iter, mode=mode, **kw)
    dt = time.perf_counter() - t0
    unit = "cycles" if mode.startswith("multigrid") else "sweeps"
    print(f"{name:22s}: {mode:14s} {s.iterations_run:6d} {unit}, final change {s.residual_history[-1]:.2e}, "
          f"{dt * 1e3:9.1f} ms wall, {s.kernel_seconds * 1e3:9.2f} ms kernels")
print(f"max |ideal - digital| = {np.max(np.abs(sol['analog, ideal'] - sol['digital'])):.2e}")
print(f"max |noisy - digital| = {np.max(np.abs(sol['analog, 1 % mismatch'] - sol['digital'])):.2e}")
em = EnergyModel(resistor_ohms=10_000, supply_voltage=1.0)
ea, ed = runs[0][1].energy_joules(em), runs[2][1].energy_joules(em)
print(f"analog {ea * 1e12:.3f} pJ, digital {ed * 1e12:.3f} pJ, gain {em.efficiency_ratio(ea, ed):.1f}x")
if not a.no_maps:
    print()
    print(ascii_heatmap(sol["digital"], "Digital solution  (ground truth)"))
    print()
    print(ascii_heatmap(sol["analog, 1 % mismatch"], "Analog  solution  (1% resistor mismatch)"))
    print()
    print(centre_column_profile({"Digital": sol["digital"], "Analog": sol["analog, ideal"], "Noisy": sol["analog, 1 % mismatch"]}))
