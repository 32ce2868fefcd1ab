#!/usr/bin/env python3
"""BASELINE config 3: N^2 Laplace, 1 % mismatch from numpy's seeded stream, fp64 exact wavefront path,
fixed K sweeps (tol = 0; the iteration does not converge at this size, SURVEY finding 2).
Checks every residual and the final grid bit for bit against the CPU oracle, prints GLUP/s."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200")); sys.path.insert(0, ROOT)
import numpy as np
from analog_pde_solver import _backend
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--noise", type=float, default=0.01)
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--warmup-k", type=int, default=2)
a = ap.parse_args()
n, k = a.n, a.k
rng = np.random.default_rng(0)
nh = nv = None
if a.noise > 0:
    nh = 1.0 + a.noise * rng.standard_normal((n, n - 1))
    nv = 1.0 + a.noise * rng.standard_normal((n - 1, n))
g0 = np.zeros((n, n)); g0[0, :] = 1.0
g = g0.copy()
_backend.solve("exact", g0.copy(), nh, nv, None, 0.0, a.warmup_k)  # warm-up (context, module load)
t0 = time.perf_counter()
hist, it, ksec = _backend.solve("exact", g, nh, nv, None, 0.0, k)
wall = time.perf_counter() - t0
lups = (n - 2) ** 2 * k
print(f"exact {n}x{n} noise={a.noise} K={k}: kernel {ksec*1e3:.2f} ms -> {lups/ksec/1e9:.2f} GLUP/s ; end-to-end {wall*1e3:.1f} ms -> {lups/wall/1e9:.2f} GLUP/s", flush=True)
if not a.no_check:
    import oracle
    t0 = time.perf_counter()
    ru, rh, _ = oracle.gauss_seidel_mt(g0, nh, nv, None, k)
    print(f"oracle ({oracle.max_threads()} threads): {time.perf_counter()-t0:.1f} s ; bit-exact grid={g.tobytes()==ru.tobytes()} residuals={hist.tobytes()==rh.tobytes()}")
