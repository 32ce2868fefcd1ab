#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the reference does not exist on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Writes tests/golden/golden_cases.npz (inputs as arrays + the reference's outputs)
and tests/golden/golden_index.json (iters / sha256 / residuals, human-readable).

Each case records exactly what the hot path consumes and produces
(analog_pde_solver/solver.py:250-290, :329-372 upstream):
  grid0    -- grid after apply_dirichlet (solver.py:250-251)
  source   -- h^2 * source_fn (solver.py:253-259), or absent
  noise_h/noise_v -- AnalogPDESolver._noise_h/_noise_v (solver.py:207-213), or absent (digital)
  solution, hist, iters -- return value, residual_history, iterations_run
"""
import hashlib
import json
import os
import sys

import numpy as np

REF = os.environ.get("APDE_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from analog_pde_solver import AnalogPDESolver, DigitalSolver, EnergyModel  # noqa: E402
from analog_pde_solver.solver import apply_dirichlet  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def top_hot(i, j, rows, cols):
    if i == 0:
        return 1.0
    if i == rows - 1 or j == 0 or j == cols - 1:
        return 0.0
    return None


def zero_boundary(i, j, rows, cols):
    if i == 0 or i == rows - 1 or j == 0 or j == cols - 1:
        return 0.0
    return None


def sin_source(i, j, rows, cols):
    x = i / (rows - 1)
    y = j / (cols - 1)
    return -2.0 * np.pi ** 2 * np.sin(np.pi * x) * np.sin(np.pi * y)


def warm_interior(i, j, rows, cols):
    """Non-None on interior nodes too: becomes the initial guess (solver.py:150-156)."""
    if i == 0:
        return 1.0
    if i == rows - 1 or j == 0 or j == cols - 1:
        return 0.25
    return 0.5 + 0.01 * ((i * 7 + j * 3) % 11)


def partial_ring(i, j, rows, cols):
    """None on part of the edge ring: those nodes stay 0.0 and are never updated."""
    if i == 0:
        return 1.0
    if j == 0:
        return 0.5
    return None


BCS = {"top_hot": top_hot, "zero_boundary": zero_boundary, "warm_interior": warm_interior,
       "partial_ring": partial_ring}
SRCS = {"none": None, "sin_source": sin_source}

# name, kind, rows, cols, sigma, seed, bc, src, tol, max_iter
CASES = [
    ("laplace20_ideal", "analog", 20, 20, 0.0, 42, "top_hot", "none", 1e-6, 20000),
    ("laplace20_noisy", "analog", 20, 20, 0.01, 0, "top_hot", "none", 1e-6, 20000),
    ("laplace20_digital", "digital", 20, 20, 0.0, None, "top_hot", "none", 1e-6, 20000),
    ("poisson20_analog", "analog", 20, 20, 0.0, None, "zero_boundary", "sin_source", 1e-7, 10000),
    ("poisson20_digital", "digital", 20, 20, 0.0, None, "zero_boundary", "sin_source", 1e-7, 10000),
    ("poisson64_digital", "digital", 64, 64, 0.0, None, "zero_boundary", "sin_source", 1e-7, 10000),
    ("poisson64_noisy", "analog", 64, 64, 0.01, 0, "zero_boundary", "sin_source", 1e-7, 10000),
    ("laplace10_tol5", "analog", 10, 10, 0.0, None, "top_hot", "none", 1e-5, 10000),
    ("laplace10_tol6", "digital", 10, 10, 0.0, None, "top_hot", "none", 1e-6, 10000),
    ("laplace10_noisy7", "analog", 10, 10, 0.01, 7, "top_hot", "none", 1e-6, 10000),
    ("poisson15", "analog", 15, 15, 0.0, None, "zero_boundary", "sin_source", 1e-7, 10000),
    ("laplace15_tol7", "digital", 15, 15, 0.0, None, "top_hot", "none", 1e-7, 10000),
    ("rect12x30_noisy_src", "analog", 12, 30, 0.01, 3, "top_hot", "sin_source", 1e-6, 10000),
    # edge cases: fixed sweep count (tol=0), max_iter cap, rows<3,
    # interior initial guess, partial ring, 5% mismatch, grids past the one-CTA kernel
    ("rect30x12_tol0_k25", "analog", 30, 12, 0.02, 11, "top_hot", "sin_source", 0.0, 25),
    ("cap_hit", "digital", 24, 24, 0.0, None, "top_hot", "none", 1e-12, 40),
    ("rows2", "analog", 2, 9, 0.01, 1, "top_hot", "none", 1e-6, 100),
    ("cols1", "digital", 5, 1, 0.0, None, "top_hot", "none", 1e-6, 100),
    ("three_by_three", "analog", 3, 3, 0.01, 5, "top_hot", "sin_source", 1e-9, 100),
    ("warm_interior16", "analog", 16, 16, 0.01, 2, "warm_interior", "none", 1e-6, 10000),
    ("partial_ring14", "digital", 14, 14, 0.0, None, "partial_ring", "sin_source", 1e-6, 10000),
    ("noisy5pct_18", "analog", 18, 18, 0.05, 9, "top_hot", "none", 1e-6, 10000),
    ("mid96x80_k12", "analog", 96, 80, 0.01, 4, "top_hot", "sin_source", 0.0, 12),
    ("mid150_k6_digital", "digital", 150, 150, 0.0, None, "top_hot", "sin_source", 0.0, 6),
    ("mid130x170_k5", "analog", 130, 170, 0.01, 6, "warm_interior", "sin_source", 0.0, 5),
]


def run_case(kind, rows, cols, sigma, seed, bc, src, tol, max_iter):
    bfn, sfn = BCS[bc], SRCS[src]
    if kind == "analog":
        s = AnalogPDESolver(rows=rows, cols=cols, noise_level=sigma, rng_seed=seed)
    else:
        s = DigitalSolver(rows=rows, cols=cols)
    u = s.solve(bfn, source_fn=sfn, tol=tol, max_iter=max_iter)
    grid0 = apply_dirichlet(np.zeros((rows, cols)), bfn)
    out = {"grid0": grid0, "solution": np.array(u), "hist": np.array(s.residual_history, dtype=np.float64),
           "iters": np.int64(s.iterations_run), "tol": np.float64(tol), "max_iter": np.int64(max_iter),
           "energy": np.float64(s.energy_joules())}
    if sfn is not None:
        h = 1.0 / (max(rows, cols) - 1)
        source = np.zeros((rows, cols))
        for i in range(rows):
            for j in range(cols):
                source[i, j] = h ** 2 * sfn(i, j, rows, cols)
        out["source"] = source
    if kind == "analog":
        out["noise_h"] = np.array(s._noise_h)
        out["noise_v"] = np.array(s._noise_v)
    return out


def main():
    blob, index = {}, {}
    for name, kind, rows, cols, sigma, seed, bc, src, tol, max_iter in CASES:
        r = run_case(kind, rows, cols, sigma, seed, bc, src, tol, max_iter)
        for k, v in r.items():
            blob[f"{name}/{k}"] = v
        hist = r["hist"]
        index[name] = {
            "kind": kind, "rows": rows, "cols": cols, "noise_level": sigma, "rng_seed": seed,
            "boundary_fn": bc, "source_fn": src, "tol": tol, "max_iter": max_iter,
            "iters": int(r["iters"]),
            "hist_first": float(hist[0]) if hist.size else None,
            "hist_last": float(hist[-1]) if hist.size else None,
            "centre": float(r["solution"][rows // 2, cols // 2]),
            "sha256_16": hashlib.sha256(r["solution"].tobytes()).hexdigest()[:16],
            "energy_joules": float(r["energy"]),
        }
        print(f"{name:24s} iters={index[name]['iters']:5d} sha={index[name]['sha256_16']}", flush=True)
    np.savez_compressed(os.path.join(HERE, "golden_cases.npz"), **blob)
    with open(os.path.join(HERE, "golden_index.json"), "w") as f:
        json.dump({"numpy": np.__version__, "reference": "danieleschmidt/analog-pde-solver-sim 1.0.0",
                   "cases": index}, f, indent=1)


if __name__ == "__main__":
    main()
