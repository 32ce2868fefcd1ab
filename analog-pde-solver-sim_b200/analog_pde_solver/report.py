"""Text reports for large grids: the reference's demo prints one character pair per node and one profile line per
row (examples/laplace_heat_demo.py:90-110 upstream), which is unreadable beyond ~40 x 40.  These helpers show the
same two views for any size by block-averaging the grid down to a terminal-sized raster first."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

SHADES = " .:-=+*#%@"   # the reference's ten shades, dark to bright (laplace_heat_demo.py:91)


def downsample(grid: np.ndarray, max_rows: int = 32, max_cols: int = 64) -> np.ndarray:
    """Block means of `grid` on a raster of at most max_rows x max_cols cells (grids that already fit are returned
    as they are).  Every node contributes to exactly one cell."""
    g = np.asarray(grid, dtype=np.float64)
    rows, cols = g.shape
    out_r, out_c = min(rows, max_rows), min(cols, max_cols)
    if (out_r, out_c) == (rows, cols):
        return g
    r_edges = np.linspace(0, rows, out_r + 1).round().astype(int)
    c_edges = np.linspace(0, cols, out_c + 1).round().astype(int)
    row_sums = np.add.reduceat(g, r_edges[:-1], axis=0)
    sums = np.add.reduceat(row_sums, c_edges[:-1], axis=1)
    counts = np.outer(np.diff(r_edges), np.diff(c_edges))
    return sums / counts


def ascii_heatmap(grid: np.ndarray, title: str = "", max_rows: int = 32, max_cols: int = 64,
                  lo: Optional[float] = None, hi: Optional[float] = None) -> str:
    """The reference's heat-map (value in [0, 1] -> one of ten shades, two characters per cell) of the
    block-averaged grid.  lo/hi rescale other value ranges; by default [0, 1] like upstream."""
    small = downsample(grid, max_rows, max_cols)
    lo = 0.0 if lo is None else lo
    hi = 1.0 if hi is None else hi
    span = (hi - lo) or 1.0
    idx = np.clip(((small - lo) / span * (len(SHADES) - 1)).astype(int), 0, len(SHADES) - 1)
    idx = np.where(np.isfinite(small), idx, 0)
    lines = [title + (f"   [{grid.shape[0]} x {grid.shape[1]} nodes shown as {small.shape[0]} x {small.shape[1]} block means]"
                      if small.shape != tuple(grid.shape) else "")]
    lines += ["".join(SHADES[k] * 2 for k in row) for row in idx]
    return "\n".join(lines)


def centre_column_profile(solutions: Dict[str, np.ndarray], max_rows: int = 24) -> str:
    """The reference's centre-column table (laplace_heat_demo.py:106-110) with at most max_rows evenly spaced rows."""
    names = list(solutions)
    first = np.asarray(solutions[names[0]])
    rows, cols = first.shape
    mid = cols // 2
    pick = np.unique(np.linspace(0, rows - 1, min(rows, max_rows)).round().astype(int))
    lines = [f"Centre-column profile (col={mid}" + (f", {len(pick)} of {rows} rows" if len(pick) < rows else "") + ")",
             "  " + f"{'Row':>7}" + "".join(f"  {n[:12]:>12}" for n in names)]
    for i in pick:
        lines.append("  " + f"{i:7d}" + "".join(f"  {np.asarray(solutions[n])[i, mid]:12.6f}" for n in names))
    return "\n".join(lines)
