"""GPU: the exact wavefront path (K1 one-CTA kernel, K2 ring kernel) is bit-identical to the
reference -- golden vectors of the unmodified reference and the pinned CPU oracle -- through the C-ABI."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN_NAMES
from helpers import BCS, SRCS, random_problem, top_hot, zero_boundary, sin_source
from analog_pde_solver import AnalogPDESolver, DigitalSolver, EnergyModel, _backend

pytestmark = pytest.mark.gpu


def _force_ring_kernel(path, monkeypatch):
    """'ring-k2' / 'ring-k2b' / 'ring-band<rows per band>' -> env for the kernel choice, returns the exact_path."""
    if not path.startswith("ring-"):
        return path
    kern = path[5:]
    if kern.startswith("band"):  # band<threads>[x<rows per thread>]; 64 threads: several bands and slots on small grids
        lpf = kern.endswith("p")  # ...p: links requested one step ahead into registers
        nt, _, r = kern[4:].rstrip("p").partition("x")
        monkeypatch.setenv("APDE_BAND_NT", nt)
        monkeypatch.setenv("APDE_BAND_R", r or "1")
        monkeypatch.setenv("APDE_BAND_LPF", "1" if lpf else "0")
        kern = "band"
    monkeypatch.setenv("APDE_RING_KERNEL", kern)
    return "ring"


def _solve(c, path):
    g = c["grid0"].copy()
    hist, n, ksec = _backend.solve("exact", g, c["noise_h"], c["noise_v"], c["source"], c["tol"],
                                   c["max_iter"], exact_path=path)
    return g, hist, n


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_bit_exact_auto_path(golden, name):
    c = golden.case(name)
    g, hist, n = _solve(c, "auto")
    assert n == c["iters"]
    assert hist.tobytes() == c["hist"].tobytes()
    assert g.tobytes() == c["solution"].tobytes()


@pytest.mark.parametrize("kern", ["default", "ring-k2", "ring-k2d", "ring-band64"])
@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if n not in ("rows2", "cols1")])
def test_golden_bit_exact_ring_path(golden, name, kern, monkeypatch):
    """Force the multi-CTA kernels on every case, including the ones that fit one CTA;
    the early-stop cases exercise snapshot / rollback / replay."""
    c = golden.case(name)
    g, hist, n = _solve(c, "ring" if kern == "default" else _force_ring_kernel(kern, monkeypatch))
    assert n == c["iters"]
    assert hist.tobytes() == c["hist"].tobytes()
    assert g.tobytes() == c["solution"].tobytes()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_python_classes_reproduce_reference(golden, name):
    """Through the drop-in classes with the reference's own callables: same iterations_run,
    residual_history, solution bytes and energy (upstream solver.py:281-295, :370-386)."""
    c = golden.case(name)
    m = c["meta"]
    if m["kind"] == "analog":
        s = AnalogPDESolver(rows=m["rows"], cols=m["cols"], noise_level=m["noise_level"], rng_seed=m["rng_seed"])
    else:
        s = DigitalSolver(rows=m["rows"], cols=m["cols"])
    u = s.solve(BCS[m["boundary_fn"]], source_fn=SRCS[m["source_fn"]], tol=m["tol"], max_iter=m["max_iter"])
    assert u is s.solution
    assert s.iterations_run == c["iters"]
    assert len(s.residual_history) == max(c["iters"], 0)
    assert np.array(s.residual_history, dtype=np.float64).tobytes() == c["hist"].tobytes()
    assert u.tobytes() == c["solution"].tobytes()
    assert s.energy_joules() == c["energy"]
    if m["kind"] == "digital":
        assert s.wall_time_s > 0


@pytest.mark.parametrize("rows,cols,sigma,src,k", [
    (3, 3, 0.01, True, 5), (3, 40, 0.02, True, 9), (40, 3, 0.02, False, 9), (4, 4, 0.0, False, 7),
    (33, 65, 0.01, True, 40), (65, 33, 0.0, True, 40), (70, 70, 0.03, False, 31),
    (257, 129, 0.01, True, 11), (300, 300, 0.0, False, 13), (513, 700, 0.01, True, 6),
])
@pytest.mark.parametrize("path", ["auto", "ring", "ring-k2", "ring-k2b", "ring-k2d", "ring-band64", "ring-band64x2", "ring-band64x3",
                                  "ring-band256x2", "ring-band256x4", "ring-band512", "ring-band64p", "ring-band64x2p",
                                  "ring-band512p"])
def test_fixed_sweeps_vs_oracle(rows, cols, sigma, src, k, path, monkeypatch):
    path = _force_ring_kernel(path, monkeypatch)  # every ring kernel on every shape (default picks one per problem)
    g0, nh, nv, s = random_problem(rows, cols, sigma, seed=rows * 1000 + cols, with_source=src)
    ru, rh, rn = oracle.gauss_seidel(g0, nh, nv, s, 0.0, k)
    g = g0.copy()
    hist, n, _ = _backend.solve("exact", g, nh, nv, s, 0.0, k, exact_path=path)
    assert n == k == rn
    assert hist.tobytes() == rh.tobytes()
    assert g.tobytes() == ru.tobytes()


@pytest.mark.parametrize("path", ["small", "ring", "ring-k2", "ring-k2b", "ring-k2d", "ring-band64", "ring-band64x3", "ring-band64x2p"])
@pytest.mark.parametrize("tol", [1e-3, 1e-5, 3e-7])
def test_early_stop_at_first_crossing(path, tol, monkeypatch):
    path = _force_ring_kernel(path, monkeypatch)
    """Stops at the true first sweep with max_change < tol although later sweeps were already in
    flight (snapshot + replay)."""
    g0, nh, nv, s = random_problem(40, 36, 0.01, seed=77)
    ru, rh, rn = oracle.gauss_seidel(g0, nh, nv, s, tol, 20000)
    g = g0.copy()
    hist, n, _ = _backend.solve("exact", g, nh, nv, s, tol, 20000, exact_path=path)
    assert n == rn and hist.tobytes() == rh.tobytes() and g.tobytes() == ru.tobytes()


@pytest.mark.parametrize("rows,cols,nt,r,kb,k", [
    (200, 130, 64, 1, 4, 150),    # 4 bands, 8 sweep slots: 19 rounds through the ring of slots
    (131, 90, 64, 1, 1, 77),      # progress exchanged every step; last band 1 row short of full
    (67, 300, 64, 1, 8, 90),      # 2 bands, the second with a single row (its thread is first AND last row)
    (600, 257, 256, 2, 4, 61),    # 256 threads x 2 rows: 2 bands, the second 86 rows high
    (1100, 520, 512, 1, 4, 45),   # 3 bands of 512, two CTAs per SM
    (300, 140, 64, 2, 4, 120),    # two rows per thread: 3 bands of 128 rows, the last 42 rows high (second walker idle)
    (260, 200, 64, 3, 2, 100),    # three rows per thread: 2 bands of 192 rows, the last row is walker 1's thread 1
    (1500, 300, 256, 4, 4, 40),   # the benchmark shape: 256 threads x 4 rows, 2 bands, the second 474 rows high
    (1030, 260, 256, 4, 4, 50),   # 256 threads x 4 rows: 2 bands, the last with 4 rows
])
@pytest.mark.parametrize("sigma,src,lpf", [(0.01, True, 0), (0.0, False, 0), (0.01, True, 1)])
def test_band_kernel_ring_of_slots_wraps(rows, cols, nt, r, kb, k, sigma, src, lpf, monkeypatch):
    """K2c with more sweeps than sweep slots: every CTA runs several tasks back to back, threads of one CTA are on
    different sweeps at the same time, and slot 0 follows slot Gs-1 of the previous round."""
    monkeypatch.setenv("APDE_RING_KERNEL", "band")
    monkeypatch.setenv("APDE_BAND_NT", str(nt))
    if lpf and (nt, r) not in ((64, 1), (64, 2), (512, 1)):
        pytest.skip("shape not instantiated with links requested a step ahead")
    monkeypatch.setenv("APDE_BAND_R", str(r))
    monkeypatch.setenv("APDE_BAND_LPF", str(lpf))
    monkeypatch.setenv("APDE_BAND_KB", str(kb))
    g0, nh, nv, s = random_problem(rows, cols, sigma, seed=rows + cols, with_source=src)
    ru, rh, rn = oracle.gauss_seidel_mt(g0, nh, nv, s, k)
    g = g0.copy()
    hist, n, _ = _backend.solve("exact", g, nh, nv, s, 0.0, k, exact_path="ring")
    assert n == k and hist.tobytes() == rh.tobytes() and g.tobytes() == ru.tobytes()


def test_ring_kernel_medium_grid_with_stop():
    g0 = np.zeros((160, 200))
    g0[0, :] = 1.0
    ru, rh, rn = oracle.gauss_seidel(g0, None, None, None, 2e-4, 5000)
    g = g0.copy()
    hist, n, _ = _backend.solve("exact", g, None, None, None, 2e-4, 5000)
    assert 100 < rn < 5000
    assert n == rn and hist.tobytes() == rh.tobytes() and g.tobytes() == ru.tobytes()


def test_large_grid_mismatch_iterates():
    """BASELINE config 3 shape at a size the oracle finishes in seconds: 1024^2 Laplace, 1 % mismatch
    from numpy's seeded stream uploaded from the host, fixed K sweeps (the iteration does not
    converge at this size, SURVEY finding 2)."""
    n, k = 1024, 24
    s = AnalogPDESolver(rows=n, cols=n, noise_level=0.01, rng_seed=0)
    g0 = np.zeros((n, n))
    g0[0, :] = 1.0
    ru, rh, _ = oracle.gauss_seidel_mt(g0, s._noise_h, s._noise_v, None, k)
    b = np.full((n, n), np.nan)
    b[0, :] = 1.0
    u = s.solve(b, tol=0.0, max_iter=k)
    assert s.iterations_run == k
    assert np.array(s.residual_history).tobytes() == rh.tobytes()
    assert u.tobytes() == ru.tobytes()


def test_nan_and_inf_inputs_follow_reference_semantics():
    """NaN changes never enter max_change (strict '>', solver.py:277) but NaN is stored (:279)."""
    g0, nh, nv, s = random_problem(12, 12, 0.01, seed=5)
    g0[5, 5] = np.nan
    ru, rh, rn = oracle.gauss_seidel(g0, nh, nv, s, 1e-6, 50)
    g = g0.copy()
    hist, n, _ = _backend.solve("exact", g, nh, nv, s, 1e-6, 50)
    assert n == rn and hist.tobytes() == rh.tobytes() and g.tobytes() == ru.tobytes()


def test_max_iter_zero_and_degenerate_shapes():
    g = np.ones((5, 5))
    hist, n, _ = _backend.solve("exact", g, None, None, None, 1e-6, 0)
    assert n == 0 and hist.size == 0 and (g == 1).all()
    d = DigitalSolver(rows=2, cols=7)
    d.solve(top_hot, tol=0.0, max_iter=4)
    assert d.iterations_run == 4 and d.residual_history == [0.0] * 4


# --- the reference's own property tests (tests/test_solver.py upstream), run on the GPU backend ---
def test_reference_properties_laplace():
    for cls in (AnalogPDESolver, DigitalSolver):
        s = cls(rows=10, cols=10)
        u = s.solve(top_hot, tol=1e-6)
        assert s.iterations_run < 10_000
        assert np.isclose(u[0, :], 1.0).all() and np.isclose(u[-1, :], 0.0).all()
        assert np.isclose(u[1:, 0], 0.0).all() and np.isclose(u[1:, -1], 0.0).all()
        assert u.min() >= -1e-8 and u.max() <= 1.0 + 1e-8
        mid = u[:, 5]
        assert all(mid[i] >= mid[i + 1] - 1e-6 for i in range(9))
        assert np.allclose(u, u[:, ::-1], atol=1e-5)
        assert s.residual_history[0] > s.residual_history[-1]
        assert s.energy_joules() > 0


def test_reference_properties_poisson_and_cross():
    exact = np.array([[np.sin(np.pi * i / 14) * np.sin(np.pi * j / 14) for j in range(15)] for i in range(15)])
    a = AnalogPDESolver(rows=15, cols=15)
    d = DigitalSolver(rows=15, cols=15)
    ua = a.solve(zero_boundary, source_fn=sin_source, tol=1e-7)
    ud = d.solve(zero_boundary, source_fn=sin_source, tol=1e-7)
    assert np.max(np.abs(ua - exact)) < 0.05 and np.max(np.abs(ud - exact)) < 0.05
    assert np.max(np.abs(ua - ud)) < 1e-4 and abs(a.iterations_run - d.iterations_run) <= 2
    noisy = AnalogPDESolver(rows=10, cols=10, noise_level=0.01, rng_seed=7).solve(top_hot, tol=1e-6)
    ideal = AnalogPDESolver(rows=10, cols=10).solve(top_hot, tol=1e-6)
    assert np.max(np.abs(noisy - ideal)) < 0.05
    a20 = AnalogPDESolver(rows=20, cols=20)
    d20 = DigitalSolver(rows=20, cols=20)
    a20.solve(top_hot, tol=1e-6, max_iter=20000)
    d20.solve(top_hot, tol=1e-6, max_iter=20000)
    em = EnergyModel()
    assert a20.energy_joules(em) < d20.energy_joules(em)
    assert f"{a20.energy_joules(em) * 1e12:.3f}" == "28.026"
    assert f"{d20.energy_joules(em) * 1e12:.3f}" == "1209600.000"
    assert f"{em.efficiency_ratio(a20.energy_joules(em), d20.energy_joules(em)):.1f}" == "43159.2"


def test_link_division_fast_path_is_ieee_exact():
    """The exact path divides by link values with a precomputed-reciprocal FMA sequence; the device
    self-test compares it with the general IEEE division on 2^28 random + adversarial operand pairs
    (all-ones / near-power-of-two significands, zeros, subnormals, huge, NaN/Inf)."""
    import ctypes
    lib = _backend.load()
    for seed in (1, 2, 3):
        bad = ctypes.c_int64(-1)
        _backend.check(lib.apde_selftest_division(1 << 28, seed, ctypes.byref(bad)))
        assert bad.value == 0


def test_unsafe_links_take_the_general_division():
    """Links outside the screened range (negative, tiny, huge, all-ones significand, zero -> inf/nan
    results) still reproduce the reference bit for bit."""
    rng = np.random.default_rng(3)
    rows, cols = 24, 20
    g0, nh, nv, s = random_problem(rows, cols, 0.3, seed=8)
    nh[3, 4] = -0.75
    nh[5, 6] = 1e-200
    nv[7, 8] = 1e200
    nv[2, 2] = np.nextafter(2.0, 0.0)       # significand all ones
    nh[9, 9] = np.nextafter(1.0, 0.0)
    g0[10, 10] = -0.0
    g0[11, 3] = 5e-324
    import os
    for path, kern in (("small", None), ("ring", "k2"), ("ring", "k2d"), ("ring", "band")):
        if kern:
            os.environ["APDE_RING_KERNEL"] = kern
        try:
            ru, rh, rn = oracle.gauss_seidel(g0, nh, nv, s, 0.0, 12)
            g = g0.copy()
            hist, n, _ = _backend.solve("exact", g, nh, nv, s, 0.0, 12, exact_path=path)
        finally:
            os.environ.pop("APDE_RING_KERNEL", None)
        assert g.tobytes() == ru.tobytes() and hist.tobytes() == rh.tobytes(), (path, kern)


def test_config3_size_4096_mismatch_iterates():
    """BASELINE config 3 at full size: 4096^2 Laplace, 1 % mismatch from numpy's seeded stream, K = 16
    sweeps; every residual and the whole grid equal the (multi-threaded, bitwise-serial) CPU oracle."""
    n, k = 4096, 16
    s = AnalogPDESolver(rows=n, cols=n, noise_level=0.01, rng_seed=0)
    g0 = np.zeros((n, n))
    g0[0, :] = 1.0
    ru, rh, _ = oracle.gauss_seidel_mt(g0, s._noise_h, s._noise_v, None, k)
    g = g0.copy()
    hist, it, _ = _backend.solve("exact", g, s._noise_h, s._noise_v, None, 0.0, k)
    assert it == k and hist.tobytes() == rh.tobytes() and g.tobytes() == ru.tobytes()


def test_config3_size_4096_many_sweeps_in_flight(monkeypatch):
    """Config 3 with more sweeps than the ring has CTA slots (148 for K2d, 2 x 148 = 296 for K2 on B200): K = 420
    sweeps, so the ring wraps (CTA c runs sweeps c, c+G, ...) with every slot busy.  Grid and all 420 residuals
    equal the CPU oracle (column-block pipelined over the host threads, bitwise equal to the serial loop) for the
    default kernel choice and for both ring kernels forced."""
    n, k = 4096, 420
    s = AnalogPDESolver(rows=n, cols=n, noise_level=0.01, rng_seed=0)
    g0 = np.zeros((n, n))
    g0[0, :] = 1.0
    ru, rh, _ = oracle.gauss_seidel_mt(g0, s._noise_h, s._noise_v, None, k)
    for kern in (None, "ring-k2", "ring-k2d"):
        if kern:
            _force_ring_kernel(kern, monkeypatch)
        g = g0.copy()
        hist, it, _ = _backend.solve("exact", g, s._noise_h, s._noise_v, None, 0.0, k)
        assert it == k and hist.tobytes() == rh.tobytes() and g.tobytes() == ru.tobytes(), kern


def test_poisson_demo_numbers():
    """What examples/poisson_demo.py upstream prints (20 x 20) and BASELINE config 2 (64 x 64)."""
    from helpers import zero_boundary, sin_source
    for n, iters, err in ((20, 458, "2.2625e-03"), (64, 4070, "1.6709e-04")):
        x = np.arange(n) / (n - 1)
        exact = np.array([[np.sin(np.pi * i / (n - 1)) * np.sin(np.pi * j / (n - 1)) for j in range(n)] for i in range(n)])
        a, d = AnalogPDESolver(rows=n, cols=n), DigitalSolver(rows=n, cols=n)
        ua = a.solve(zero_boundary, source_fn=sin_source, tol=1e-7)
        ud = d.solve(zero_boundary, source_fn=sin_source, tol=1e-7)
        assert a.iterations_run == d.iterations_run == iters
        assert f"{np.max(np.abs(ua - exact)):.4e}" == err and f"{np.max(np.abs(ud - exact)):.4e}" == err
        if n == 20:
            em = EnergyModel()
            ea, ed = a.energy_joules(em), d.energy_joules(em)
            assert f"{ea * 1e12:.3f}" == "24.730" and f"{ed * 1e12:.3f}" == "1648800.000"
            assert f"{em.efficiency_ratio(ea, ed):.1f}" == "66671.7"
