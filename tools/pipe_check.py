#!/usr/bin/env python3
"""Quick correctness probe of one red-black kernel family against the CPU oracle (development aid; the real
parity tests are tests/test_redblack_gpu.py).  usage: APDE_RB_VARIANT=v2 python tools/pipe_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from helpers import random_problem
from analog_pde_solver import _backend

bad = 0
for (rows, cols) in [(5, 4), (17, 33), (64, 64), (130, 67), (200, 1031), (700, 1500)]:
    for sigma, src in [(0.0, True), (0.0, False), (0.02, True)]:
        for mode, dt in [("redblack", np.float64), ("redblack32", np.float32)]:
            for depth in [1, 2, 3, 4, 5, 6]:
                for rb in ("40", ""):
                    if rb:
                        os.environ["APDE_ROW_BLOCK"] = rb
                    else:
                        os.environ.pop("APDE_ROW_BLOCK", None)
                    k = 13
                    g0, nh, nv, s = random_problem(rows, cols, sigma, seed=rows * 131 + cols, with_source=src)
                    ru, rh, _ = oracle.red_black(g0, nh, nv, s, k, dtype=dt)
                    g = g0.copy()
                    hist, n, _ = _backend.solve(mode, g, nh, nv, s, 0.0, k, check_every=4, temporal_depth=depth)
                    ok_g = g.astype(dt).tobytes() == ru.tobytes()
                    ok_h = hist.astype(dt).tobytes() == rh.tobytes()
                    if not (ok_g and ok_h):
                        bad += 1
                        d = np.abs(g.astype(dt).astype(np.float64) - ru.astype(np.float64))
                        ij = np.unravel_index(np.argmax(d), d.shape)
                        nbad = int((d > 0).sum())
                        rr, cc = np.nonzero(d > 0)
                        print(f"MISMATCH {rows}x{cols} sigma={sigma} src={src} {mode} depth={depth} rowblk={rb or 'auto'}: grid={ok_g} hist={ok_h} "
                              f"nbad={nbad} max={d.max():.3e} at {ij} rows[{rr.min() if nbad else -1},{rr.max() if nbad else -1}] cols[{cc.min() if nbad else -1},{cc.max() if nbad else -1}]", flush=True)
                        if bad > 12:
                            sys.exit(1)
print("PIPE_CHECK", "PASS" if bad == 0 else f"FAIL ({bad})")
sys.exit(0 if bad == 0 else 1)
