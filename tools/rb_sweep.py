#!/usr/bin/env python3
"""Parameter sweep of the red-black streaming kernel on one GPU (tuning aid, not the bench contract).
usage: python tools/rb_sweep.py [--n 16384] [--dtypes f64,f32] [--depths 1,2,3,4] [--lanes 1,2] [--noise 0,1]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200"))
import numpy as np
from analog_pde_solver import _backend

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--dtypes", default="f64,f32")
ap.add_argument("--depths", default="1,2,3,4")
ap.add_argument("--lanes", default="default")
ap.add_argument("--noise", default="0")
ap.add_argument("--src", default="1")
ap.add_argument("--rowblocks", default="auto")
ap.add_argument("--sweeps", type=int, default=24)
a = ap.parse_args()
n = a.n
for dt in a.dtypes.split(","):
    for noise in [int(x) for x in a.noise.split(",")]:
        for src in [int(x) for x in a.src.split(",")]:
            for lv in a.lanes.split(","):
                for rb in a.rowblocks.split(","):
                    if lv == "default":
                        os.environ.pop("APDE_LANE_VECTORS", None)
                    else:
                        os.environ["APDE_LANE_VECTORS"] = lv
                    if rb == "auto":
                        os.environ.pop("APDE_ROW_BLOCK", None)
                    else:
                        os.environ["APDE_ROW_BLOCK"] = rb
                    for depth in [int(x) for x in a.depths.split(",")]:
                        p = _backend.Plan(n, n, dtype=np.float64 if dt == "f64" else np.float32,
                                          has_noise=bool(noise), has_source=bool(src), temporal_depth=depth)
                        p.fill(kind=1 if src else 0, top_hot=not src, noise_level=0.01 if noise else 0.0, seed=7)
                        sweeps = (a.sweeps // depth) * depth
                        p.run(sweeps, 0); p.residuals(0, 1)
                        best = 1e9
                        for rep in range(3):
                            t0 = time.perf_counter()
                            p.run(sweeps, 0)
                            p.residuals(0, 1)
                            best = min(best, time.perf_counter() - t0)
                        glups = (n - 2) ** 2 * sweeps / best / 1e9
                        print(f"{dt} noise={noise} src={src} lanes={lv} rowblk={rb} depth={depth}: {glups:8.1f} GLUP/s  ({best/sweeps*1e3:.3f} ms/sweep)", flush=True)
                        p.close()
