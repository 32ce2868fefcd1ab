"""CPU: the C-ABI library loads and exports what include/apde.h declares; the Python mirror keeps
the reference's surface; nothing in the product imports the oracle; no silent CPU fallback."""
import inspect
import os
import re
import sys

import numpy as np
import pytest

from conftest import HAVE_GPU, PKG_DIR, ROOT
from analog_pde_solver import AnalogPDESolver, DigitalSolver, EnergyModel, _backend
import analog_pde_solver


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "apde.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(apde_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _backend.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for sym in declared:
        assert hasattr(lib, sym), f"libapde.so does not export {sym}"
    assert sorted(_backend.EXPORTED_SYMBOLS) == declared
    assert lib.apde_version() == 100


def test_python_surface_matches_reference():
    assert analog_pde_solver.__all__ == ["AnalogPDESolver", "DigitalSolver", "EnergyModel"]
    assert analog_pde_solver.__version__ == "1.0.0"
    sig = inspect.signature(AnalogPDESolver.__init__)
    assert [(p.name, p.default) for p in list(sig.parameters.values())[1:]] == [
        ("rows", 20), ("cols", 20), ("resistor_ohms", 10_000.0), ("noise_level", 0.0), ("rng_seed", None)]
    for cls in (AnalogPDESolver, DigitalSolver):
        ps = list(inspect.signature(cls.solve).parameters.values())[1:]
        pos = [(p.name, p.default) for p in ps if p.kind == p.POSITIONAL_OR_KEYWORD]
        assert pos == [("boundary_fn", inspect.Parameter.empty), ("source_fn", None), ("tol", 1e-6), ("max_iter", 10_000)]
        assert all(p.kind == p.KEYWORD_ONLY for p in ps[4:])
    assert [p.name for p in list(inspect.signature(DigitalSolver.__init__).parameters.values())[1:]] == ["rows", "cols"]
    assert [p.name for p in list(inspect.signature(DigitalSolver.energy_joules).parameters.values())[1:]] == [
        "energy_model", "flops_per_node_per_iter"]
    s = AnalogPDESolver(rows=6, cols=7, noise_level=0.01, rng_seed=3)
    assert s._noise_h.shape == (6, 6) and s._noise_v.shape == (5, 7)
    assert s.solution.shape == (6, 7) and s.iterations_run == 0 and s.residual_history == []
    d = DigitalSolver(rows=4, cols=5)
    assert d.wall_time_s == 0.0 and d.solution.shape == (4, 5)


def test_mismatch_arrays_follow_numpy_stream_order(golden):
    """_noise_h is drawn before _noise_v from default_rng(seed) (solver.py:207-213 upstream)."""
    for name in ("laplace20_noisy", "rect12x30_noisy_src", "laplace10_noisy7"):
        c = golden.case(name)
        m = c["meta"]
        s = AnalogPDESolver(rows=m["rows"], cols=m["cols"], noise_level=m["noise_level"], rng_seed=m["rng_seed"])
        assert s._noise_h.tobytes() == c["noise_h"].tobytes()
        assert s._noise_v.tobytes() == c["noise_v"].tobytes()
    ideal = AnalogPDESolver(rows=5, cols=4)
    assert (ideal._noise_h == 1.0).all() and (ideal._noise_v == 1.0).all()


def test_host_setup_matches_reference_arrays(golden):
    """apply_dirichlet + source build reproduce grid0/source of the reference bit for bit."""
    from analog_pde_solver.solver import _initial_grid, _source_term
    from helpers import BCS, SRCS
    for name in golden.names():
        c = golden.case(name)
        m = c["meta"]
        g = _initial_grid(m["rows"], m["cols"], BCS[m["boundary_fn"]])
        assert g.tobytes() == c["grid0"].tobytes(), name
        s = _source_term(m["rows"], m["cols"], SRCS[m["source_fn"]])
        if c["source"] is None:
            assert s is None
        else:
            assert s.tobytes() == c["source"].tobytes(), name


def test_array_inputs_equal_callable_inputs():
    from analog_pde_solver.solver import _initial_grid, _source_term
    from helpers import top_hot, sin_source
    rows, cols = 9, 13
    b = np.full((rows, cols), np.nan)
    b[0, :] = 1.0
    b[-1, :] = 0.0
    b[1:, 0] = 0.0
    b[1:, -1] = 0.0
    assert _initial_grid(rows, cols, b).tobytes() == _initial_grid(rows, cols, top_hot).tobytes()
    f = np.array([[sin_source(i, j, rows, cols) for j in range(cols)] for i in range(rows)])
    assert _source_term(rows, cols, f).tobytes() == _source_term(rows, cols, sin_source).tobytes()


def test_energy_model_matches_reference_values(golden):
    """EnergyModel arithmetic on the reference's own solutions gives the reference's joules."""
    for name in golden.names():
        c = golden.case(name)
        m = c["meta"]
        if m["kind"] == "analog":
            s = AnalogPDESolver(rows=m["rows"], cols=m["cols"])
            s.solution = c["solution"]
            assert s.energy_joules() == c["energy"], name
        else:
            d = DigitalSolver(rows=m["rows"], cols=m["cols"])
            d.iterations_run = c["iters"]
            assert d.energy_joules() == c["energy"], name
    em = EnergyModel()
    assert em.efficiency_ratio(0.0, 1.0) == float("inf")
    assert em.digital_energy_joules(10, 10, 200) == 2 * em.digital_energy_joules(10, 10, 100)
    # README numbers: 20x20 Laplace, 336 iterations -> 1 209 600 pJ digital, 28.026 pJ analog
    assert round(em.digital_energy_joules(20, 20, 336) * 1e12, 3) == 1209600.0
    assert round(golden.case("laplace20_ideal")["energy"] * 1e12, 3) == 28.026


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(PKG_DIR):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, fn
                assert "apde_oracle" not in text, fn


@pytest.mark.skipif(HAVE_GPU, reason="checks the failure mode without a GPU")
def test_no_cpu_fallback_without_gpu():
    from helpers import top_hot
    with pytest.raises(_backend.ApdeError) as e:
        DigitalSolver(rows=8, cols=8).solve(top_hot)
    assert e.value.code == -2
    with pytest.raises(_backend.ApdeError):
        AnalogPDESolver(rows=8, cols=8).solve(top_hot, mode="redblack")
    with pytest.raises(_backend.ApdeError):
        _backend.Plan(64, 64)


def test_bad_arguments_are_rejected_before_any_device_work():
    lib = _backend.load()
    import ctypes
    n = ctypes.c_int64(0)
    rc = lib.apde_solve_exact_f64(None, None, None, None, 4, 4, 1e-6, 10, None, ctypes.byref(n), None)
    assert rc == -1 and b"bad argument" in lib.apde_last_error()
    g = np.zeros((4, 4))
    with pytest.raises(ValueError):
        _backend.solve("exact", g, np.ones((4, 4)), None, None, 1e-6, 10)
    with pytest.raises(ValueError):
        _backend.solve("nonsense", g, None, None, None, 1e-6, 10)
    # the windowed fixed-sweep solve: wrong mode / shapes in Python, too few rows or a bad window in the library -- all before
    # any device work (this container has no GPU)
    with pytest.raises(ValueError):
        _backend.solve_window("exact", np.zeros((4000, 64)), None, None, None, 4, 0, 100)
    with pytest.raises(ValueError):
        _backend.solve_window("redblack", np.zeros((4000, 64)), np.ones((4000, 64)), None, None, 4, 0, 100)
    for args in ((np.zeros((100, 64)), 4, 0, 50), (np.zeros((4000, 64)), 4, 200, 100), (np.zeros((4000, 64)), 4, 0, 4001)):
        with pytest.raises(_backend.ApdeError) as e:
            _backend.solve_window("redblack", args[0], None, None, None, *args[1:])
        assert e.value.code == -1


def test_degenerate_grid_raises_like_the_reference():
    """A 1 x 1 grid divides by zero when the reference evaluates h = 1/(max(rows, cols) - 1) (solver.py:254, :329),
    with or without a source; the mirror evaluates h first too.  No GPU is touched before the error."""
    from analog_pde_solver import AnalogPDESolver, DigitalSolver
    for solver in (AnalogPDESolver(rows=1, cols=1), DigitalSolver(rows=1, cols=1)):
        with pytest.raises(ZeroDivisionError):
            solver.solve(lambda i, j, r, c: 0.0)


def test_environment_options_are_validated(monkeypatch):
    """Malformed APDE_* values fail with a message naming the variable, not deep inside solve()."""
    from analog_pde_solver import DigitalSolver
    d = DigitalSolver(rows=8, cols=8)
    for name, value in (("APDE_MODE", "fastest"), ("APDE_CHECK_EVERY", "often"), ("APDE_GPUS", "0"),
                        ("APDE_EXACT_PATH", "diagonal"), ("APDE_OMEGA", "big")):
        monkeypatch.setenv(name, value)
        with pytest.raises(ValueError, match=name if name != "APDE_MODE" else "fastest"):
            d.solve(lambda i, j, r, c: 1.0 if i == 0 else None, mode=None if name == "APDE_MODE" else "redblack")
        monkeypatch.delenv(name)


def test_band_kernel_flag_protocol_never_deadlocks():
    """The exact path's row-band kernel (K2c) is a cooperative kernel whose CTAs wait for each other: a launch shape
    that can deadlock would hang the GPU.  tools/band_protocol_sim.py replays the flag protocol on the CPU with the
    launch shape the host picks (slots * (3*kb + 4) <= cols): nothing may be left unfinished -- and the first version
    of the protocol (progress withheld while a CTA polls) is shown to deadlock on the shape that exposed it."""
    import random
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import band_protocol_sim as sim
    rng = random.Random(7)
    for _ in range(60):
        nt, kb = rng.choice([64, 128, 192, 256, 512, 1024]), rng.choice([1, 2, 4, 8, 16])  # nt = band height
        rows, cols, k = rng.randint(3, 1200), rng.randint(3 * kb + 4, 400), rng.randint(1, 80)
        assert sim.replay(rows, cols, nt, kb, k) == [], (rows, cols, nt, kb, k)
    for shape in ((67, 300, 64, 8, 90), (200, 130, 64, 4, 150), (131, 90, 64, 1, 77), (1100, 520, 512, 4, 45)):
        assert sim.replay(*shape) == [], shape
    assert sim.replay(67, 300, 64, 8, 90, publish_first=False, slots=15) != []
