import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200"))
os.environ["APDE_TRACE"] = "1"
import numpy as np, torch
from analog_pde_solver import _backend
n = 16384
pin = lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
u = pin((n, n)); u[:] = 0.0
s = pin((n, n)); s[:] = 1e-9
for rep in range(3):
    t0 = time.perf_counter()
    _backend.solve("redblack", u, None, None, s, 0.0, 256, check_every=32, temporal_depth=3)
    print("call", rep, time.perf_counter() - t0, flush=True)
up = np.zeros((n, n)); sp = np.full((n, n), 1e-9)
t0 = time.perf_counter()
_backend.solve("redblack", up, None, None, sp, 0.0, 256, check_every=32, temporal_depth=3)
print("pageable call", time.perf_counter() - t0, flush=True)
