"""Resistive-mesh PDE solvers with the relaxation sweeps on a B200 GPU.

Host-side mirror of the reference module ``analog_pde_solver/solver.py`` (upstream): the same
three public classes, constructor arguments, ``solve()`` signatures, result attributes and
energy accounting -- but the sweep loops (upstream solver.py:262-287 and :344-368) are ONE call
into ``libapde.so`` instead of Python loops.

Backend selection never changes a signature.  It comes from, in order: the keyword-only
``mode=`` argument of ``solve()``, the instance attribute ``backend_mode``, the environment
variable ``APDE_MODE``; default ``"exact"``:

* ``exact``      -- fp64 wavefront Gauss-Seidel, bit-identical to the reference on every iterate
                    (same ``iterations_run``, ``residual_history``, ``solution`` bytes).
* ``redblack``   -- fp64 red-black throughput path: same fixed point, different update order, so
                    ``iterations_run`` may differ by a few sweeps; convergence is tested every
                    ``APDE_CHECK_EVERY`` (default 32) sweeps.
* ``redblack32`` -- the same in fp32.
* ``redblack_sor`` / ``redblack_sor32`` -- red-black with over-relaxation (keyword-only ``omega=``, default the
                    optimum for the ideal mesh): O(N) instead of O(N^2) sweeps, same fixed point.
* ``redblack_kcl`` / ``redblack_kcl32`` -- each node divides by the sum of its own link conductances (Kirchhoff's
                    current law) instead of the reference's constant 4: convergent at every grid size where the
                    reference's update diverges with mismatch; a DIFFERENT (physically normalised) fixed point
                    when links are mismatched, the same one when they are ideal.  ``omega=`` applies too.

* ``multigrid`` / ``multigrid32`` -- geometric V-cycles with the red-black kernel as smoother, any grid size; with
                    mismatched links it solves the normalised (``redblack_kcl``) system: ``iterations_run`` counts CYCLES (each about 5 fine sweeps of work) and
                    ``residual_history`` holds one entry per cycle (how far the cycle moved any node); ``max_iter``
                    caps the cycles.  16384 x 16384 Poisson reaches its discretisation error in 12 cycles.

``APDE_GPUS=N`` spreads a red-black solve over N GPUs of the box from this one process (row strips,
CUDA peer copies); the exact path is sequential along anti-diagonals and always uses one GPU.

``boundary_fn`` / ``source_fn`` may also be ``numpy`` arrays (extension for large grids where a
per-node Python callable costs minutes): a boundary array is the initial grid itself (NaN =
"None", i.e. left at 0.0), a source array holds f(i, j) and is scaled by h**2 like upstream
solver.py:254-259.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import Callable, Optional, Union

import numpy as np

from . import _backend

BoundaryFn = Callable[[int, int, int, int], Optional[float]]
SourceFn = Callable[[int, int, int, int], float]

_NODE_CAPACITANCE_F = 1e-12  # 1 pF parasitic node capacitance (upstream solver.py:81)


# ----------------------------------------------------------------------------------------------
# energy accounting -- host numpy, O(N^2) once per solve, semantics of upstream solver.py:48-127
# ----------------------------------------------------------------------------------------------
@dataclass
class EnergyModel:
    """Analog-vs-digital energy estimate (upstream solver.py:48-127).

    resistor_ohms       mesh resistance; supply_voltage is accepted and unused, as upstream.
    flop_energy_*_pj    energy per digital add / multiply in pJ.
    """

    resistor_ohms: float = 10_000.0
    supply_voltage: float = 1.0
    flop_energy_add_pj: float = 1.0
    flop_energy_mul_pj: float = 5.0

    def analog_energy_joules(self, solution: np.ndarray, settling_time_s: Optional[float] = None) -> float:
        """sum(dV^2)/R over horizontal then vertical links, times 5*R*C settling (solver.py:70-97)."""
        drop_h = np.diff(solution, axis=1)
        drop_v = np.diff(solution, axis=0)
        watts = np.sum(drop_h ** 2) / self.resistor_ohms + np.sum(drop_v ** 2) / self.resistor_ohms
        if settling_time_s is None:
            settling_time_s = 5.0 * self.resistor_ohms * _NODE_CAPACITANCE_F
        return watts * settling_time_s

    def digital_energy_joules(self, grid_rows: int, grid_cols: int, iterations: int,
                              flops_per_node_per_iter: int = 5) -> float:
        """(4 adds + 1 mul) per node per sweep over ALL rows*cols nodes (solver.py:99-117).

        ``flops_per_node_per_iter`` is accepted and ignored, as upstream.
        """
        node_sweeps = grid_rows * grid_cols * iterations
        pj = 4 * node_sweeps * self.flop_energy_add_pj + 1 * node_sweeps * self.flop_energy_mul_pj
        return pj * 1e-12

    def efficiency_ratio(self, analog_J: float, digital_J: float) -> float:
        """digital / analog, inf when the analog energy is zero (solver.py:119-127)."""
        return float("inf") if analog_J == 0 else digital_J / analog_J


# ----------------------------------------------------------------------------------------------
# host setup shared by both solvers (upstream solver.py:134-156, :250-259, :329-339)
# ----------------------------------------------------------------------------------------------
def apply_dirichlet(grid: np.ndarray, boundary_fn: BoundaryFn) -> np.ndarray:
    """Write ``boundary_fn(i, j, rows, cols)`` wherever it is not None (any node, in place)."""
    rows, cols = grid.shape
    for i in range(rows):
        row = grid[i]
        for j in range(cols):
            value = boundary_fn(i, j, rows, cols)
            if value is not None:
                row[j] = value
    return grid


def _initial_grid(rows: int, cols: int, boundary: Union[BoundaryFn, np.ndarray]) -> np.ndarray:
    grid = np.zeros((rows, cols))
    if isinstance(boundary, np.ndarray):
        if boundary.shape != (rows, cols):
            raise ValueError(f"boundary array must have shape {(rows, cols)}, got {boundary.shape}")
        b = np.asarray(boundary, dtype=np.float64)
        keep = ~np.isnan(b)
        grid[keep] = b[keep]
        return grid
    return apply_dirichlet(grid, boundary)


def _source_term(rows: int, cols: int, source: Union[None, SourceFn, np.ndarray]) -> Optional[np.ndarray]:
    """h^2 * f with h = 1/(max(rows, cols) - 1); None for the Laplace equation.

    ``h`` is evaluated first, as upstream does (solver.py:254, :329): a 1 x 1 grid raises
    ZeroDivisionError there even without a source, so it does here.
    """
    h = 1.0 / (max(rows, cols) - 1)
    if source is None:
        return None
    if isinstance(source, np.ndarray):
        if source.shape != (rows, cols):
            raise ValueError(f"source array must have shape {(rows, cols)}, got {source.shape}")
        return h ** 2 * np.asarray(source, dtype=np.float64)
    term = np.zeros((rows, cols))
    scale = h ** 2
    for i in range(rows):
        row = term[i]
        for j in range(cols):
            row[j] = scale * source(i, j, rows, cols)
    return term


def _resolve_mode(explicit: Optional[str], instance_default: Optional[str]) -> str:
    mode = (explicit or instance_default or os.environ.get("APDE_MODE") or "exact").lower()
    if mode not in _backend.MODES:
        raise ValueError(f"unknown APDE mode {mode!r}: expected one of {sorted(_backend.MODES)}")
    return mode


def _env_int(name: str, default: int, lo: int) -> int:
    """Integer option from the environment with a message that names the variable."""
    raw = os.environ.get(name)
    if raw is None or raw.strip() == "":
        return default
    try:
        value = int(raw)
    except ValueError:
        raise ValueError(f"{name}={raw!r} is not an integer") from None
    if value < lo:
        raise ValueError(f"{name}={value} must be >= {lo}")
    return value


def _env_float(name: str, default: Optional[float]) -> Optional[float]:
    raw = os.environ.get(name)
    if raw is None or raw.strip() == "":
        return default
    try:
        return float(raw)
    except ValueError:
        raise ValueError(f"{name}={raw!r} is not a number") from None


def _relax(solver, grid, noise_h, noise_v, source, tol, max_iter, mode, omega=None):
    """The one device call that replaces the reference's sweep loop."""
    exact_path = os.environ.get("APDE_EXACT_PATH", "auto").lower()
    if exact_path not in _backend.EXACT_PATHS:
        raise ValueError(f"APDE_EXACT_PATH={exact_path!r}: expected one of {sorted(_backend.EXACT_PATHS)}")
    hist, iters, ksec = _backend.solve(
        mode, grid, noise_h, noise_v, source, tol, max_iter,
        check_every=_env_int("APDE_CHECK_EVERY", 0, 0),
        temporal_depth=_env_int("APDE_TEMPORAL_DEPTH", 0, 0),
        exact_path=exact_path,
        n_gpus=_env_int("APDE_GPUS", 1, 1),
        omega=omega if omega is not None else _env_float("APDE_OMEGA", None))
    solver.residual_history = hist.tolist()
    solver.iterations_run = iters
    solver.kernel_seconds = ksec
    solver.solution = grid
    return grid


# ----------------------------------------------------------------------------------------------
# AnalogPDESolver -- upstream solver.py:162-295
# ----------------------------------------------------------------------------------------------
class AnalogPDESolver:
    """Resistive-mesh analog computer for the Laplace / Poisson equation, relaxed on the GPU.

    rows, cols       grid size including the boundary ring
    resistor_ohms    link resistance (energy estimate only)
    noise_level      relative resistor mismatch sigma (0 = ideal, 0.01 = 1 %)
    rng_seed         seed of numpy's default_rng for the mismatch arrays

    The mismatch arrays are drawn on the host exactly as upstream (solver.py:201-213: ``_noise_h``
    of shape (rows, cols-1) first, then ``_noise_v`` of shape (rows-1, cols)) and uploaded, so a
    seeded run sees the same resistors as the reference.
    """

    backend_mode: Optional[str] = None

    def __init__(self, rows: int = 20, cols: int = 20, resistor_ohms: float = 10_000.0,
                 noise_level: float = 0.0, rng_seed: Optional[int] = None) -> None:
        self.rows = rows
        self.cols = cols
        self.resistor_ohms = resistor_ohms
        self.noise_level = noise_level
        self._rng = np.random.default_rng(rng_seed)
        self.solution: np.ndarray = np.zeros((rows, cols))
        self.iterations_run: int = 0
        self.residual_history: list[float] = []
        self.kernel_seconds: float = 0.0
        if noise_level > 0:
            self._noise_h = 1.0 + noise_level * self._rng.standard_normal((rows, cols - 1))
            self._noise_v = 1.0 + noise_level * self._rng.standard_normal((rows - 1, cols))
        else:
            self._noise_h = np.ones((rows, cols - 1))
            self._noise_v = np.ones((rows - 1, cols))

    def solve(self, boundary_fn: Union[BoundaryFn, np.ndarray],
              source_fn: Union[None, SourceFn, np.ndarray] = None,
              tol: float = 1e-6, max_iter: int = 10_000, *, mode: Optional[str] = None,
              omega: Optional[float] = None) -> np.ndarray:
        """Relax to ``max_change < tol`` or ``max_iter`` sweeps; returns (and stores) the grid."""
        rows, cols = self.rows, self.cols
        grid = _initial_grid(rows, cols, boundary_fn)
        source = _source_term(rows, cols, source_fn)
        # all-ones links divide exactly (x / 1.0 == x): skip them, the result has the same bits.  The
        # arrays are checked, not the flag, so links edited after construction are honoured like upstream.
        ideal = (not (self.noise_level > 0)) and bool((self._noise_h == 1.0).all() and (self._noise_v == 1.0).all())
        return _relax(self, grid, None if ideal else self._noise_h, None if ideal else self._noise_v,
                      source, tol, max_iter, _resolve_mode(mode, self.backend_mode), omega)

    def energy_joules(self, energy_model: Optional[EnergyModel] = None) -> float:
        model = energy_model or EnergyModel(resistor_ohms=self.resistor_ohms)
        return model.analog_energy_joules(self.solution)


# ----------------------------------------------------------------------------------------------
# DigitalSolver -- upstream solver.py:302-386
# ----------------------------------------------------------------------------------------------
class DigitalSolver:
    """Finite-difference Gauss-Seidel for the same stencil, relaxed on the GPU.

    ``energy_joules()`` charges ``iterations_run`` sweeps like upstream (solver.py:381-386).  In the default
    ``exact`` mode that count is the reference's; the red-black modes test convergence every ``APDE_CHECK_EVERY``
    (32) sweeps, so their count -- and the digital energy derived from it -- is rounded up to the end of the check
    interval (at most 31 sweeps more than the first crossing); ``multigrid`` counts cycles, not sweeps.
    """

    backend_mode: Optional[str] = None

    def __init__(self, rows: int = 20, cols: int = 20) -> None:
        self.rows = rows
        self.cols = cols
        self.solution: np.ndarray = np.zeros((rows, cols))
        self.iterations_run: int = 0
        self.residual_history: list[float] = []
        self.wall_time_s: float = 0.0
        self.kernel_seconds: float = 0.0

    def solve(self, boundary_fn: Union[BoundaryFn, np.ndarray],
              source_fn: Union[None, SourceFn, np.ndarray] = None,
              tol: float = 1e-6, max_iter: int = 10_000, *, mode: Optional[str] = None,
              omega: Optional[float] = None) -> np.ndarray:
        """Same contract as :meth:`AnalogPDESolver.solve`; also records ``wall_time_s``.

        ``wall_time_s`` brackets the relaxation only (host->device copy, kernels, device->host copy),
        not the boundary/source setup -- the window upstream measures (solver.py:341, :370).
        """
        rows, cols = self.rows, self.cols
        grid = _initial_grid(rows, cols, boundary_fn)
        source = _source_term(rows, cols, source_fn)
        started = time.perf_counter()
        _relax(self, grid, None, None, source, tol, max_iter, _resolve_mode(mode, self.backend_mode), omega)
        self.wall_time_s = time.perf_counter() - started
        return grid

    def energy_joules(self, energy_model: Optional[EnergyModel] = None,
                      flops_per_node_per_iter: int = 5) -> float:
        model = energy_model or EnergyModel()
        return model.digital_energy_joules(self.rows, self.cols, self.iterations_run, flops_per_node_per_iter)
