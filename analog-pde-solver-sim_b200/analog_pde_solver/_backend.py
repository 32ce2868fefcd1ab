"""ctypes binding of libapde.so (C-ABI declared in include/apde.h).

The library is the only compute path: if it is missing, or no sm_100 GPU is usable, every solve
raises.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("APDE_LIBRARY", os.path.join(_ROOT, "lib", "libapde.so"))

_dp = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
_i64p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p

APDE_OK = 0
ERROR_NAMES = {-1: "APDE_EINVAL", -2: "APDE_ENODEV", -3: "APDE_ECUDA", -4: "APDE_ENOMEM", -5: "APDE_ESTATE"}

# every symbol include/apde.h declares (tests check that the shared object exports each one)
EXPORTED_SYMBOLS = (
    "apde_version", "apde_device_count", "apde_last_error",
    "apde_solve_exact_f64", "apde_solve_exact_f64_path", "apde_solve_redblack_f64", "apde_solve_redblack_f32", "apde_solve_redblack_multi", "apde_release_cache",
    "apde_plan_create", "apde_plan_destroy", "apde_plan_upload", "apde_plan_upload_rows", "apde_plan_download_rows", "apde_plan_fill",
    "apde_plan_download", "apde_plan_clear_residuals", "apde_plan_run", "apde_plan_halo", "apde_plan_residuals",
    "apde_plan_read_residuals", "apde_plan_launch_count", "apde_plan_checksum", "apde_plan_checksum_bits", "apde_memcpy_dtod", "apde_selftest_division", "apde_plan_energy_sums", "apde_plan_download_array", "apde_default_temporal_depth", "apde_solve_redblack_ex", "apde_plan_set_relaxation",
    "apde_solve_multigrid", "apde_multigrid_levels", "apde_solve_redblack_window",
)


class ApdeError(RuntimeError):
    """A libapde call failed; carries the C error code."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class PlanDesc(ctypes.Structure):
    _fields_ = [("dtype_bytes", ctypes.c_int32), ("has_noise", ctypes.c_int32),
                ("has_source", ctypes.c_int32), ("temporal_depth", ctypes.c_int32),
                ("grows", _i64), ("cols", _i64), ("row_begin", _i64), ("row_end", _i64)]


class FillDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("top_hot", ctypes.c_int32),
                ("noise_level", ctypes.c_double), ("seed", ctypes.c_uint64)]


_lib = None


def load():
    """Load libapde.so once.  Raises (never falls back) when the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  analog_pde_solver has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.apde_version.restype = ctypes.c_int
    lib.apde_device_count.restype = ctypes.c_int
    lib.apde_last_error.restype = ctypes.c_char_p
    solve_tail = [_i64, _i64, ctypes.c_double, _i64]
    lib.apde_solve_exact_f64.restype = ctypes.c_int
    lib.apde_solve_exact_f64.argtypes = [_dp, _dp, _dp, _dp] + solve_tail + [_dp, _i64p, _dp]
    lib.apde_solve_exact_f64_path.restype = ctypes.c_int
    lib.apde_solve_exact_f64_path.argtypes = [_dp, _dp, _dp, _dp] + solve_tail + [_dp, _i64p, _dp, ctypes.c_int]
    for name in ("apde_solve_redblack_f64", "apde_solve_redblack_f32"):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [_dp, _dp, _dp, _dp] + solve_tail + [ctypes.c_int, ctypes.c_int, _dp, _i64p, _dp]
    lib.apde_solve_redblack_multi.restype = ctypes.c_int
    lib.apde_solve_redblack_multi.argtypes = [ctypes.c_int, _dp, _dp, _dp, _dp] + solve_tail + [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _i64p, _dp]
    lib.apde_solve_redblack_window.restype = ctypes.c_int
    lib.apde_solve_redblack_window.argtypes = [ctypes.c_int, _dp, _dp, _dp, _dp, _i64, _i64, _i64, ctypes.c_int, _i64, _i64, _dp, _dp]
    lib.apde_plan_create.restype = ctypes.c_int
    lib.apde_plan_create.argtypes = [ctypes.POINTER(_vp), ctypes.POINTER(PlanDesc)]
    lib.apde_plan_destroy.restype = ctypes.c_int
    lib.apde_plan_destroy.argtypes = [_vp]
    lib.apde_plan_upload.restype = ctypes.c_int
    lib.apde_plan_upload.argtypes = [_vp, _dp, _dp, _dp, _dp]
    lib.apde_plan_upload_rows.restype = ctypes.c_int
    lib.apde_plan_upload_rows.argtypes = [_vp, _i64, _i64, _dp, _dp, _dp, _dp]
    lib.apde_plan_download_rows.restype = ctypes.c_int
    lib.apde_plan_download_rows.argtypes = [_vp, _i64, _i64, _dp]
    lib.apde_solve_redblack_ex.restype = ctypes.c_int
    lib.apde_solve_redblack_ex.argtypes = [ctypes.c_int, _dp, _dp, _dp, _dp] + solve_tail + [
        ctypes.c_int, ctypes.c_double, ctypes.c_int, _dp, _i64p, _dp]
    lib.apde_solve_multigrid.restype = ctypes.c_int
    lib.apde_solve_multigrid.argtypes = [ctypes.c_int, _dp, _dp, _dp, _dp, _i64, _i64, ctypes.c_double, _i64, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, _dp, _i64p, _dp]
    lib.apde_multigrid_levels.restype = ctypes.c_int
    lib.apde_multigrid_levels.argtypes = [_i64, _i64]
    lib.apde_plan_set_relaxation.restype = ctypes.c_int
    lib.apde_plan_set_relaxation.argtypes = [_vp, ctypes.c_double, ctypes.c_int]
    lib.apde_default_temporal_depth.restype = ctypes.c_int
    lib.apde_default_temporal_depth.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.apde_plan_download_array.restype = ctypes.c_int
    lib.apde_plan_download_array.argtypes = [_vp, ctypes.c_int, _i64, _i64, _dp]
    lib.apde_plan_fill.restype = ctypes.c_int
    lib.apde_plan_fill.argtypes = [_vp, ctypes.POINTER(FillDesc)]
    lib.apde_plan_download.restype = ctypes.c_int
    lib.apde_plan_download.argtypes = [_vp, _dp]
    lib.apde_plan_run.restype = ctypes.c_int
    lib.apde_plan_run.argtypes = [_vp, _i64, _i64, ctypes.c_int, _vp]
    lib.apde_plan_clear_residuals.restype = ctypes.c_int
    lib.apde_plan_clear_residuals.argtypes = [_vp, _i64, _i64, _vp]
    lib.apde_plan_halo.restype = ctypes.c_int
    lib.apde_plan_halo.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64p]
    lib.apde_plan_residuals.restype = ctypes.c_int
    lib.apde_plan_residuals.argtypes = [_vp, ctypes.POINTER(_vp), _i64p]
    lib.apde_plan_read_residuals.restype = ctypes.c_int
    lib.apde_plan_read_residuals.argtypes = [_vp, _i64, _i64, _dp, _vp]
    lib.apde_plan_launch_count.restype = _i64
    lib.apde_plan_launch_count.argtypes = [_vp]
    lib.apde_plan_checksum.restype = ctypes.c_int
    lib.apde_plan_checksum.argtypes = [_vp, _dp, _dp, _vp]
    lib.apde_plan_checksum_bits.restype = ctypes.c_int
    lib.apde_plan_checksum_bits.argtypes = [_vp, ctypes.POINTER(ctypes.c_uint64), _vp]
    lib.apde_plan_energy_sums.restype = ctypes.c_int
    lib.apde_plan_energy_sums.argtypes = [_vp, _dp, _dp, _vp]
    lib.apde_selftest_division.restype = ctypes.c_int
    lib.apde_selftest_division.argtypes = [_i64, ctypes.c_uint64, _i64p]
    lib.apde_memcpy_dtod.restype = ctypes.c_int
    lib.apde_memcpy_dtod.argtypes = [_vp, _vp, _i64, _vp]
    _lib = lib
    return lib


def default_temporal_depth(dtype: str, has_noise: bool) -> int:
    """Fused sweeps per HBM pass the library picks for this variant ("f64" / "f32")."""
    return int(load().apde_default_temporal_depth(8 if dtype in ("f64", "float64") else 4, int(bool(has_noise))))


def multigrid_levels(rows: int, cols: int) -> int:
    """Number of grid levels a V-cycle gets on a rows x cols grid (1 = too small to coarsen)."""
    return int(load().apde_multigrid_levels(int(rows), int(cols)))


def check(rc: int) -> None:
    if rc != APDE_OK:
        msg = load().apde_last_error()
        raise ApdeError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")


def _ptr(a: Optional[np.ndarray]):
    return ctypes.cast(None, _dp) if a is None else a.ctypes.data_as(_dp)


def _as_f64(a, shape, name):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
    return a


EXACT_PATHS = {"auto": 0, "small": 1, "ring": 2}
MODES = ("exact", "redblack", "redblack64", "redblack32",
         "redblack_sor", "redblack_sor32", "redblack_kcl", "redblack_kcl32", "multigrid", "multigrid32")
EXTENDED_MODES = {  # mode -> (dtype bytes, normalised conductances, over-relaxed by default)
    "redblack_sor": (8, False, True), "redblack_sor32": (4, False, True),
    "redblack_kcl": (8, True, False), "redblack_kcl32": (4, True, False),
}


def optimal_omega(rows: int, cols: int) -> float:
    """Optimal SOR factor of the 5-point Laplacian on a rows x cols Dirichlet grid:
    2 / (1 + sqrt(1 - rho^2)) with the Jacobi spectral radius rho = (cos(pi/(rows-1)) + cos(pi/(cols-1))) / 2
    (= 2 / (1 + sin(pi/(N-1))) on a square grid)."""
    if rows < 3 or cols < 3:
        return 1.0
    rho = 0.5 * (np.cos(np.pi / (rows - 1)) + np.cos(np.pi / (cols - 1)))
    return float(2.0 / (1.0 + np.sqrt(max(1.0 - rho * rho, 0.0))))

# Sweeps per C call.  The reference accepts any max_iter because its residual list grows lazily
# (solver.py:262, :281); here the history buffer of one call is max_iter doubles on the host and on the
# device, so a "practically unlimited" max_iter is served in chunks that continue from the grid the
# previous chunk left (Gauss-Seidel's only state), which is bit-identical to one long call.
MAX_SWEEPS_PER_CALL = 1 << 20


def _solve_once(lib, mode, grid, nh, nv, src, tol, max_iter, check_every, temporal_depth, exact_path, n_gpus,
                omega=None):
    rows, cols = grid.shape
    hist = np.zeros(max(max_iter, 1), dtype=np.float64)
    iters = ctypes.c_int64(0)
    ksec = ctypes.c_double(0.0)
    head = (_ptr(grid), _ptr(nh), _ptr(nv), _ptr(src), rows, cols, float(tol), max_iter)
    if mode == "exact":
        rc = lib.apde_solve_exact_f64_path(*head, _ptr(hist), ctypes.byref(iters), ctypes.byref(ksec),
                                           EXACT_PATHS[exact_path])
    elif mode in ("multigrid", "multigrid32"):
        nu = int(os.environ.get("APDE_MG_SMOOTH", "0"))
        rc = lib.apde_solve_multigrid(4 if mode == "multigrid32" else 8, _ptr(grid), _ptr(nh), _ptr(nv), _ptr(src), rows, cols, float(tol),
                                      max_iter, nu, nu, int(os.environ.get("APDE_MG_COARSE_SWEEPS", "0")), _ptr(hist),
                                      ctypes.byref(iters), ctypes.byref(ksec))
    elif mode in EXTENDED_MODES:
        elem, kcl, sor = EXTENDED_MODES[mode]
        if omega is None:
            omega = optimal_omega(rows, cols) if sor else 1.0
        if not (0.0 < float(omega) < 2.0):
            raise ValueError(f"omega must lie in (0, 2), got {omega}")
        rc = lib.apde_solve_redblack_ex(elem, *head, int(check_every), float(omega), int(kcl), _ptr(hist),
                                        ctypes.byref(iters), ctypes.byref(ksec))
    elif n_gpus > 1:
        rc = lib.apde_solve_redblack_multi(4 if mode == "redblack32" else 8, *head, int(check_every),
                                           int(temporal_depth), int(n_gpus), _ptr(hist), ctypes.byref(iters),
                                           ctypes.byref(ksec))
    elif mode in ("redblack", "redblack64"):
        rc = lib.apde_solve_redblack_f64(*head, int(check_every), int(temporal_depth), _ptr(hist),
                                         ctypes.byref(iters), ctypes.byref(ksec))
    else:
        rc = lib.apde_solve_redblack_f32(*head, int(check_every), int(temporal_depth), _ptr(hist),
                                         ctypes.byref(iters), ctypes.byref(ksec))
    check(rc)
    n = int(iters.value)
    return hist[:max(n, 0)], n, float(ksec.value)


def solve_window(mode: str, grid: np.ndarray, noise_h, noise_v, source, n_sweeps: int, win_lo: int, win_hi: int,
                 temporal_depth: int = 0):
    """apde_solve_redblack_window: exactly n_sweeps red-black sweeps of `grid` (in place, rows [win_lo, win_hi) only are
    written back), residual history over the window rows.  Returns (hist, kernel_seconds)."""
    lib = load()
    if mode not in ("redblack", "redblack64", "redblack32"):
        raise ValueError(f"solve_window: mode must be redblack or redblack32, got {mode!r}")
    if grid.dtype != np.float64 or not grid.flags.c_contiguous or grid.ndim != 2:
        raise ValueError("grid must be a C-contiguous float64 2-D array")
    rows, cols = grid.shape
    nh = _as_f64(noise_h, (rows, cols - 1), "noise_h")
    nv = _as_f64(noise_v, (rows - 1, cols), "noise_v")
    src = _as_f64(source, (rows, cols), "source")
    hist = np.zeros(max(int(n_sweeps), 1), dtype=np.float64)
    ksec = ctypes.c_double(0.0)
    check(lib.apde_solve_redblack_window(4 if mode == "redblack32" else 8, _ptr(grid), _ptr(nh), _ptr(nv), _ptr(src), rows, cols,
                                         int(n_sweeps), int(temporal_depth), int(win_lo), int(win_hi), _ptr(hist),
                                         ctypes.byref(ksec)))
    return hist[:max(int(n_sweeps), 0)], float(ksec.value)


def solve(mode: str, grid: np.ndarray, noise_h, noise_v, source, tol: float, max_iter: int,
          check_every: int = 0, temporal_depth: int = 0, exact_path: str = "auto", n_gpus: int = 1,
          omega: Optional[float] = None):
    """One relaxation solve on the GPU.  `grid` (float64, C-order) is updated in place.

    mode: "exact" (bit-exact wavefront, fp64), "redblack" (fp64), "redblack32" (fp32), or the extended updates
    "redblack_sor[32]" (over-relaxation, `omega` defaults to optimal_omega(rows, cols)) and "redblack_kcl[32]"
    (normalised conductances; `omega` defaults to 1).
    Returns (residual_history ndarray, iterations_run, kernel_seconds).
    """
    lib = load()
    if grid.dtype != np.float64 or not grid.flags.c_contiguous or grid.ndim != 2:
        raise ValueError("grid must be a C-contiguous float64 2-D array")
    if mode not in MODES:
        raise ValueError(f"unknown APDE mode {mode!r} (expected one of {', '.join(MODES)})")
    if exact_path not in EXACT_PATHS:
        raise ValueError(f"unknown exact path {exact_path!r} (expected one of {', '.join(EXACT_PATHS)})")
    rows, cols = grid.shape
    nh = _as_f64(noise_h, (rows, cols - 1), "noise_h")
    nv = _as_f64(noise_v, (rows - 1, cols), "noise_v")
    src = _as_f64(source, (rows, cols), "source")
    max_iter = int(max_iter)
    if max_iter <= MAX_SWEEPS_PER_CALL:
        hist, n, ksec = _solve_once(lib, mode, grid, nh, nv, src, tol, max_iter, check_every, temporal_depth,
                                    exact_path, n_gpus, omega)
        return hist.copy(), n, ksec
    parts, done, ksec_total = [], 0, 0.0
    while done < max_iter:
        chunk = min(MAX_SWEEPS_PER_CALL, max_iter - done)
        hist, n, ksec = _solve_once(lib, mode, grid, nh, nv, src, tol, chunk, check_every, temporal_depth,
                                    exact_path, n_gpus, omega)
        parts.append(hist.copy())
        done += n
        ksec_total += ksec
        # the chunk stopped early, or the stopping sweep sits in its last check interval (solver.py:283-285)
        tail = 1 if mode == "exact" else (int(check_every) if check_every > 0 else 32)
        if n < chunk or (n > 0 and bool((hist[max(n - tail, 0):n] < tol).any())):
            break
    return (np.concatenate(parts) if parts else np.zeros(0)), done, ksec_total


class Plan:
    """Device-resident row strip of a global grid (apde_plan_* in include/apde.h)."""

    def __init__(self, grows: int, cols: int, row_begin: int = 0, row_end: Optional[int] = None,
                 dtype=np.float64, has_noise: bool = False, has_source: bool = False,
                 temporal_depth: int = 4):
        self._lib = load()
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("dtype must be float32 or float64")
        row_end = grows if row_end is None else row_end
        self.desc = PlanDesc(self.dtype.itemsize, int(has_noise), int(has_source), int(temporal_depth),
                             int(grows), int(cols), int(row_begin), int(row_end))
        self._h = _vp(None)
        check(self._lib.apde_plan_create(ctypes.byref(self._h), ctypes.byref(self.desc)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.apde_plan_destroy(self._h)
            self._h = _vp(None)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def upload(self, grid, noise_h=None, noise_v=None, source=None):
        g, c = self.desc.grows, self.desc.cols
        grid = _as_f64(grid, (g, c), "grid")
        check(self._lib.apde_plan_upload(self._h, _ptr(grid), _ptr(_as_f64(noise_h, (g, c - 1), "noise_h")),
                                         _ptr(_as_f64(noise_v, (g - 1, c), "noise_v")),
                                         _ptr(_as_f64(source, (g, c), "source"))))

    def fill(self, kind: int = 0, top_hot: bool = False, noise_level: float = 0.0, seed: int = 0):
        f = FillDesc(int(kind), int(top_hot), float(noise_level), int(seed))
        check(self._lib.apde_plan_fill(self._h, ctypes.byref(f)))

    def download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        g, c = self.desc.grows, self.desc.cols
        if out is None:
            out = np.zeros((g, c), dtype=np.float64)
        check(self._lib.apde_plan_download(self._h, _ptr(out)))
        return out

    ARRAYS = {"grid": 0, "source": 1, "gh": 2, "gv": 3}

    def download_array(self, which: str, row_lo: int = 0, row_hi: Optional[int] = None) -> np.ndarray:
        """Rows [row_lo, row_hi) of one of the plan's arrays as float64 (owned rows only; others stay 0)."""
        row_hi = self.desc.grows if row_hi is None else row_hi
        out = np.zeros((row_hi - row_lo, self.desc.cols), dtype=np.float64)
        check(self._lib.apde_plan_download_array(self._h, self.ARRAYS[which], int(row_lo), int(row_hi), _ptr(out)))
        return out

    def set_relaxation(self, omega: float = 1.0, normalised: bool = False):
        """Extended update for the following run() calls (over-relaxation / normalised conductances)."""
        check(self._lib.apde_plan_set_relaxation(self._h, float(omega), int(bool(normalised))))

    def run(self, n_sweeps: int, sweep_base: int = 0, phase: int = 0, stream: int = 0):
        check(self._lib.apde_plan_run(self._h, int(n_sweeps), int(sweep_base), int(phase), _vp(stream)))

    def clear_residuals(self, first: int, count: int, stream: int = 0):
        check(self._lib.apde_plan_clear_residuals(self._h, int(first), int(count), _vp(stream)))

    def halo(self, side: int):
        s, r, n = _vp(None), _vp(None), ctypes.c_int64(0)
        check(self._lib.apde_plan_halo(self._h, int(side), ctypes.byref(s), ctypes.byref(r), ctypes.byref(n)))
        return s.value, r.value, int(n.value)

    def residuals(self, first: int, count: int, stream: int = 0) -> np.ndarray:
        out = np.zeros(max(int(count), 1), dtype=np.float64)
        check(self._lib.apde_plan_read_residuals(self._h, int(first), int(count), _ptr(out), _vp(stream)))
        return out[:count]

    def checksum(self, stream: int = 0):
        s, m = ctypes.c_double(0.0), ctypes.c_double(0.0)
        check(self._lib.apde_plan_checksum(self._h, ctypes.byref(s), ctypes.byref(m), _vp(stream)))
        return float(s.value), float(m.value)

    def checksum_bits(self, stream: int = 0) -> int:
        v = ctypes.c_uint64(0)
        check(self._lib.apde_plan_checksum_bits(self._h, ctypes.byref(v), _vp(stream)))
        return int(v.value)

    def energy_sums(self, stream: int = 0):
        """(sum of squared horizontal drops, sum of squared vertical drops) over this strip's links."""
        h, v = ctypes.c_double(0.0), ctypes.c_double(0.0)
        check(self._lib.apde_plan_energy_sums(self._h, ctypes.byref(h), ctypes.byref(v), _vp(stream)))
        return float(h.value), float(v.value)

    @property
    def launches(self) -> int:
        return int(self._lib.apde_plan_launch_count(self._h))
