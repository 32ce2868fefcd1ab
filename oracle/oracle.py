"""ctypes front-end of the CPU oracle (oracle/apde_oracle.c).  TEST INFRASTRUCTURE ONLY.

Parity status: pinned -- see the header of apde_oracle.c and tests/test_oracle.py.
Every function mirrors a piece of the reference's hot path
(analog_pde_solver/solver.py:262-287, :344-368 upstream).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libapde_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (idempotent)."""
    src = os.path.join(_HERE, "apde_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libapde_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, env={**os.environ, "CC": "gcc"})
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.apde_oracle_gs_f64.restype = _i64
        lib.apde_oracle_gs_f64.argtypes = [_dp, _dp, _dp, _dp, _i64, _i64, ctypes.c_double, _i64, _dp]
        lib.apde_oracle_gs_f64_mt.restype = _i64
        lib.apde_oracle_gs_f64_mt.argtypes = [_dp, _dp, _dp, _dp, _i64, _i64, _i64, _dp, ctypes.c_int]
        lib.apde_oracle_rb_f64.restype = _i64
        lib.apde_oracle_rb_f64.argtypes = [_dp, _dp, _dp, _dp, _i64, _i64, _i64, _dp]
        lib.apde_oracle_rb_f32.restype = _i64
        lib.apde_oracle_rb_f32.argtypes = [_fp, _fp, _fp, _fp, _i64, _i64, _i64, _fp]
        lib.apde_oracle_rbx_f64.restype = _i64
        lib.apde_oracle_rbx_f64.argtypes = [_dp, _dp, _dp, _dp, _i64, _i64, _i64, ctypes.c_double, ctypes.c_int, _dp]
        lib.apde_oracle_rbx_f32.restype = _i64
        lib.apde_oracle_rbx_f32.argtypes = [_fp, _fp, _fp, _fp, _i64, _i64, _i64, ctypes.c_float, ctypes.c_int, _fp]
        lib.apde_oracle_recip_f64.restype = None
        lib.apde_oracle_recip_f64.argtypes = [_dp, _dp, _i64]
        lib.apde_oracle_recip_f32.restype = None
        lib.apde_oracle_recip_f32.argtypes = [_dp, _fp, _i64]
        lib.apde_oracle_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def _ptr(a: Optional[np.ndarray], ct):
    if a is None:
        return ctypes.cast(None, ct)
    return a.ctypes.data_as(ct)


def _chk(a, shape, dtype, name):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.shape != shape:
        raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
    return a


def max_threads() -> int:
    return int(_load().apde_oracle_max_threads())


def gauss_seidel(grid: np.ndarray, noise_h=None, noise_v=None, source=None,
                 tol: float = 1e-6, max_iter: int = 10_000) -> Tuple[np.ndarray, np.ndarray, int]:
    """Lexicographic Gauss-Seidel exactly as solver.py:262-287 / :344-368.

    `grid` holds the Dirichlet ring + initial guess (solver.py:250-251).  Returns
    (solution, residual_history, iterations_run).  The input is not modified.
    """
    lib = _load()
    u = np.array(grid, dtype=np.float64, order="C", copy=True)
    rows, cols = u.shape
    nh = _chk(noise_h, (rows, cols - 1), np.float64, "noise_h")
    nv = _chk(noise_v, (rows - 1, cols), np.float64, "noise_v")
    src = _chk(source, (rows, cols), np.float64, "source")
    hist = np.zeros(max(int(max_iter), 1), dtype=np.float64)
    n = lib.apde_oracle_gs_f64(_ptr(u, _dp), _ptr(nh, _dp), _ptr(nv, _dp), _ptr(src, _dp),
                               rows, cols, float(tol), int(max_iter), _ptr(hist, _dp))
    return u, hist[:n].copy(), int(n)


def gauss_seidel_mt(grid, noise_h=None, noise_v=None, source=None, n_sweeps: int = 1,
                    n_threads: int = 0):
    """Column-block pipelined multi-thread variant; bitwise equal to gauss_seidel(tol=0)."""
    lib = _load()
    u = np.array(grid, dtype=np.float64, order="C", copy=True)
    rows, cols = u.shape
    nh = _chk(noise_h, (rows, cols - 1), np.float64, "noise_h")
    nv = _chk(noise_v, (rows - 1, cols), np.float64, "noise_v")
    src = _chk(source, (rows, cols), np.float64, "source")
    hist = np.zeros(max(int(n_sweeps), 1), dtype=np.float64)
    if n_threads <= 0:
        n_threads = max_threads()
    lib.apde_oracle_gs_f64_mt(_ptr(u, _dp), _ptr(nh, _dp), _ptr(nv, _dp), _ptr(src, _dp),
                              rows, cols, int(n_sweeps), _ptr(hist, _dp), int(n_threads))
    return u, hist[:n_sweeps].copy(), int(n_sweeps)


def reciprocal_links(noise: np.ndarray, dtype) -> np.ndarray:
    """g = 1/noise in fp64, rounded once to `dtype` (what kernel K5 computes)."""
    lib = _load()
    noise = np.ascontiguousarray(noise, dtype=np.float64)
    out = np.empty(noise.shape, dtype=dtype)
    if np.dtype(dtype) == np.float64:
        lib.apde_oracle_recip_f64(_ptr(noise, _dp), _ptr(out, _dp), noise.size)
    else:
        lib.apde_oracle_recip_f32(_ptr(noise, _dp), _ptr(out, _fp), noise.size)
    return out


def red_black(grid, noise_h=None, noise_v=None, source=None, n_sweeps: int = 1,
              dtype=np.float64, conductances=None, omega: float = 1.0, kcl: bool = False):
    """Red-black Gauss-Seidel, colour (i+j) even first, fixed sweep count.

    The CPU statement of the CUDA throughput kernels' arithmetic (expression tree
    documented in apde_oracle.c).  Host-side inputs are fp64 like the C-ABI's; the
    working type is `dtype`.  Returns (grid in `dtype`, hist in `dtype`, n_sweeps).
    """
    lib = _load()
    dtype = np.dtype(dtype)
    u = np.array(grid, dtype=np.float64, order="C", copy=True).astype(dtype)
    rows, cols = u.shape
    if (noise_h is None) != (noise_v is None):
        raise ValueError("noise_h and noise_v must be given together")
    gh = gv = None
    if conductances is not None:
        # g given directly in the working type (e.g. read back from the device): gh (rows, cols-1), gv (rows-1, cols)
        gh = _chk(conductances[0], (rows, cols - 1), dtype, "gh")
        gv = _chk(conductances[1], (rows - 1, cols), dtype, "gv")
    elif noise_h is not None:
        gh = reciprocal_links(_chk(noise_h, (rows, cols - 1), np.float64, "noise_h"), dtype)
        gv = reciprocal_links(_chk(noise_v, (rows - 1, cols), np.float64, "noise_v"), dtype)
    src = None
    if source is not None:
        src = _chk(source, (rows, cols), np.float64, "source").astype(dtype)
    hist = np.zeros(max(int(n_sweeps), 1), dtype=dtype)
    if omega != 1.0 or kcl:
        # extended update (over-relaxation / normalised conductances), see RBX_BODY in apde_oracle.c
        if dtype == np.float64:
            lib.apde_oracle_rbx_f64(_ptr(u, _dp), _ptr(gh, _dp), _ptr(gv, _dp), _ptr(src, _dp),
                                    rows, cols, int(n_sweeps), float(omega), int(bool(kcl)), _ptr(hist, _dp))
        elif dtype == np.float32:
            lib.apde_oracle_rbx_f32(_ptr(u, _fp), _ptr(gh, _fp), _ptr(gv, _fp), _ptr(src, _fp),
                                    rows, cols, int(n_sweeps), float(np.float32(omega)), int(bool(kcl)), _ptr(hist, _fp))
        else:
            raise ValueError("dtype must be float32 or float64")
        return u, hist[:n_sweeps].copy(), int(n_sweeps)
    if dtype == np.float64:
        lib.apde_oracle_rb_f64(_ptr(u, _dp), _ptr(gh, _dp), _ptr(gv, _dp), _ptr(src, _dp),
                               rows, cols, int(n_sweeps), _ptr(hist, _dp))
    elif dtype == np.float32:
        lib.apde_oracle_rb_f32(_ptr(u, _fp), _ptr(gh, _fp), _ptr(gv, _fp), _ptr(src, _fp),
                               rows, cols, int(n_sweeps), _ptr(hist, _fp))
    else:
        raise ValueError("dtype must be float32 or float64")
    return u, hist[:n_sweeps].copy(), int(n_sweeps)


def _mg_levels(rows, cols):
    n = 1
    while rows > 3 and cols > 3 and (rows + 1) // 2 >= 3 and (cols + 1) // 2 >= 3:
        rows, cols = (rows + 1) // 2, (cols + 1) // 2
        n += 1
    return n


def _mg_hat_matrix(n_fine, n_coarse):
    """P[f, I] = max(0, 1 - |f * (1/r) - I|), r = (n_fine-1)/(n_coarse-1): the device's mg_hat, in fp64."""
    inv_r = 1.0 / ((n_fine - 1) / (n_coarse - 1))
    f = np.arange(n_fine, dtype=np.float64)[:, None]
    big_i = np.arange(n_coarse, dtype=np.float64)[None, :]
    return np.maximum(0.0, 1.0 - np.abs(f * inv_r - big_i))


def multigrid(grid, source=None, tol: float = 1e-8, max_cycles: int = 50, nu1: int = 2, nu2: int = 2,
              coarse_sweeps: int = 32, dtype=np.float64, noise_h=None, noise_v=None):
    """CPU statement of csrc/apde_multigrid.cu (NOT in the reference): V(nu1, nu2) cycles on the ideal mesh,
    red-black smoothing (red_black above), r = ((((uN+uS)+uW)+uE) - 4u) - src, levels of ((n+1)//2) nodes per side,
    hat-function transfer weights from the node positions (restriction = transpose of linear interpolation, sums over
    ascending fine indices, rows outer; prolongation rows of the coarse cell first) -- the same association as the
    device code, all transfer arithmetic in fp64, so results match the GPU bit for bit.
    Per cycle the history holds an upper bound of how far the cycle moved any fine node: the fp64 sum, in order, of the
    pre-smoothing sweeps' max_change, max |coarse-grid correction| and the post-smoothing sweeps' max_change.
    Returns (grid in dtype, that history, cycles)."""
    dt = np.dtype(dtype).type
    u = np.array(grid, dtype=np.float64).astype(dtype)
    rows, cols = u.shape
    src = None if source is None else np.asarray(source, dtype=np.float64).astype(dtype)
    L = _mg_levels(rows, cols)
    four = dt(4)
    gh = gv = None
    if noise_h is not None:   # links act on the finest level only: normalised (Kirchhoff) update and residual there
        gh = reciprocal_links(np.asarray(noise_h, dtype=np.float64), dtype)
        gv = reciprocal_links(np.asarray(noise_v, dtype=np.float64), dtype)

    def smooth(uu, ss, n, fine=False):
        if fine and gh is not None:
            out, h, _ = red_black(uu.astype(np.float64), None, None, None if ss is None else ss.astype(np.float64), n,
                                  dtype=dtype, conductances=(gh, gv), kcl=True)
        else:
            out, h, _ = red_black(uu.astype(np.float64), None, None, None if ss is None else ss.astype(np.float64), n, dtype=dtype)
        return out, h

    def restrict(uu, ss, fine=False):
        r = np.zeros(uu.shape, dtype=np.float64)
        if fine and gh is not None:
            gn, gs_, gw, ge = gv[:-1, 1:-1], gv[1:, 1:-1], gh[1:-1, :-1], gh[1:-1, 1:]
            t = gn * uu[:-2, 1:-1] + gs_ * uu[2:, 1:-1]
            t = t + gw * uu[1:-1, :-2]
            t = t + ge * uu[1:-1, 2:]
            t = t - (((gn + gs_) + gw) + ge) * uu[1:-1, 1:-1]
        else:
            t = (uu[:-2, 1:-1] + uu[2:, 1:-1])
            t = t + uu[1:-1, :-2]
            t = t + uu[1:-1, 2:]
            t = t - four * uu[1:-1, 1:-1]
        if ss is not None:
            t = t - ss[1:-1, 1:-1]
        r[1:-1, 1:-1] = t.astype(np.float64)
        nf_r, nf_c = uu.shape
        rc, cc = (nf_r + 1) // 2, (nf_c + 1) // 2
        pr, pc = _mg_hat_matrix(nf_r, rc), _mg_hat_matrix(nf_c, cc)
        # row[f, J] = sum_g (ascending, zero weights skipped) P[g, J] * r[f, g]; acc[I, J] = sum_f P[f, I] * row[f, J]
        row = np.zeros((nf_r, cc))
        for g in range(1, nf_c - 1):
            js = np.nonzero(pc[g])[0]
            for j in js:
                row[:, j] = row[:, j] + pc[g, j] * r[:, g]
        acc = np.zeros((rc, cc))
        for f in range(1, nf_r - 1):
            for i in np.nonzero(pr[f])[0]:
                acc[i, :] = acc[i, :] + pr[f, i] * row[f, :]
        sc = np.zeros((rc, cc), dtype=dtype)
        sc[1:-1, 1:-1] = (-acc[1:-1, 1:-1]).astype(dtype)
        return sc

    def prolong(uu, e, stats=None):
        nf_r, nf_c = uu.shape
        rc, cc = e.shape
        inv_rx, inv_ry = 1.0 / ((nf_r - 1) / (rc - 1)), 1.0 / ((nf_c - 1) / (cc - 1))
        xi, xj = np.arange(nf_r, dtype=np.float64) * inv_rx, np.arange(nf_c, dtype=np.float64) * inv_ry
        bi, bj = np.minimum(np.floor(xi).astype(int), rc - 2), np.minimum(np.floor(xj).astype(int), cc - 2)
        ti, tj = (xi - bi)[:, None], (xj - bj)[None, :]
        ed = e.astype(np.float64)
        e00, e01 = ed[bi][:, bj], ed[bi][:, bj + 1]
        e10, e11 = ed[bi + 1][:, bj], ed[bi + 1][:, bj + 1]
        top = (1.0 - tj) * e00 + tj * e01
        bot = (1.0 - tj) * e10 + tj * e11
        corr = ((1.0 - ti) * top + ti * bot).astype(dtype)
        out = uu.copy()
        out[1:-1, 1:-1] = uu[1:-1, 1:-1] + corr[1:-1, 1:-1]
        if stats is not None:
            a = np.abs(corr[1:-1, 1:-1])
            a = a[~np.isnan(a)]
            stats.append(float(a.max()) if a.size else 0.0)
        return out

    hist = []
    for _ in range(int(max_cycles)):
        us, ss = [u], [src]
        pre = post = np.zeros(0)
        corr = []
        for l in range(L - 1):
            us[l], h = smooth(us[l], ss[l], nu1, fine=(l == 0))
            if l == 0:
                pre = h
            sc = restrict(us[l], ss[l], fine=(l == 0))
            us.append(np.zeros_like(sc))
            ss.append(sc)
        us[L - 1], h = smooth(us[L - 1], ss[L - 1], coarse_sweeps if L > 1 else nu1 + nu2, fine=(L == 1))
        if L == 1:
            pre, post = h[:nu1], h[nu1:]
        for l in range(L - 2, -1, -1):
            us[l] = prolong(us[l], us[l + 1], corr if l == 0 else None)
            us[l], h = smooth(us[l], ss[l], nu2, fine=(l == 0))
            if l == 0:
                post = h
        u = us[0]
        moved = 0.0
        for v in pre:
            moved += float(v)
        moved += corr[0] if corr else 0.0
        for v in post:
            moved += float(v)
        hist.append(moved)
        if moved < tol:
            break
    return u, np.array(hist, dtype=np.float64), len(hist)


__all__ = ["build", "gauss_seidel", "gauss_seidel_mt", "red_black", "reciprocal_links", "max_threads", "multigrid"]
