#!/usr/bin/env python3
"""Run ONE red-black configuration a few times (profiling target for ncu).
usage: python tools/one_config.py f64 4 0 1 [n] [sweeps]   -> dtype depth noise src"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200"))
import numpy as np
from analog_pde_solver import _backend
dt, depth, noise, src = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 16384
sweeps = int(sys.argv[6]) if len(sys.argv) > 6 else 4 * depth
p = _backend.Plan(n, n, dtype=np.float64 if dt == "f64" else np.float32, has_noise=bool(noise), has_source=bool(src), temporal_depth=depth)
p.fill(kind=1 if src else 0, top_hot=not src, noise_level=0.01 if noise else 0.0, seed=7)
p.run(sweeps, 0); p.residuals(0, 1)
t0 = time.perf_counter(); p.run(sweeps, 0); p.residuals(0, 1); dtm = time.perf_counter() - t0
print(f"{dt} depth={depth} noise={noise} src={src}: {(n-2)**2*sweeps/dtm/1e9:.1f} GLUP/s")
p.close()
