"""Small end-to-end runs of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "analog-pde-solver-sim_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from helpers import random_problem
from analog_pde_solver import _backend
ok = True
for rows, cols in ((20, 20), (37, 53)):
    g0, nh, nv, s = random_problem(rows, cols, 0.01, 1)
    ru, rh, rn = oracle.gauss_seidel(g0, nh, nv, s, 1e-5, 500)
    for path in ("small", "ring"):
        for kern in ("k2", "k2b"):
            os.environ["APDE_RING_KERNEL"] = kern
            g = g0.copy()
            h, n, _ = _backend.solve("exact", g, nh, nv, s, 1e-5, 500, exact_path=path)
            ok &= (n == rn and g.tobytes() == ru.tobytes() and h.tobytes() == rh.tobytes())
for rows, cols in ((70, 300), (5, 9)):
    g0, nh, nv, s = random_problem(rows, cols, 0.01, 2)
    for mode, dt in (("redblack", np.float64), ("redblack32", np.float32)):
        for depth in (1, 2, 3, 4):
            for lv in ("1", "2"):
                os.environ["APDE_LANE_VECTORS"] = lv
                os.environ["APDE_ROW_BLOCK"] = "24"
                ru, rh, _ = oracle.red_black(g0, nh, nv, s, 5, dtype=dt)
                g = g0.copy()
                h, n, _ = _backend.solve(mode, g, nh, nv, s, 0.0, 5, check_every=5, temporal_depth=depth)
                ok &= (g.astype(dt).tobytes() == ru.tobytes() and h.astype(dt).tobytes() == rh.tobytes())
                g = g0.copy()
                h, n, _ = _backend.solve(mode, g, None, None, s, 0.0, 5, check_every=5, temporal_depth=depth)
# strips with phases
kw = dict(dtype=np.float64, has_noise=True, has_source=True, temporal_depth=2)
g0, nh, nv, s = random_problem(90, 70, 0.01, 3)
a = _backend.Plan(90, 70, 0, 40, **kw); b = _backend.Plan(90, 70, 40, 90, **kw)
lib = _backend.load()
for p in (a, b):
    p.upload(g0, nh, nv, s); p.clear_residuals(0, 2); p.run(2, 0, phase=1)
sa, ra, n = a.halo(1); sb, rb, _ = b.halo(0)
_backend.check(lib.apde_memcpy_dtod(rb, sa, n, None)); _backend.check(lib.apde_memcpy_dtod(ra, sb, n, None))
for p in (a, b):
    p.run(2, 0, phase=2); p.checksum(); p.energy_sums(); p.checksum_bits()
out = np.zeros((90, 70)); a.download(out); b.download(out); a.close(); b.close()
ref, _, _ = oracle.red_black(g0, nh, nv, s, 2)
ok &= out.tobytes() == ref.tobytes()
print("SANITIZE_SMOKE", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
