"""Boundary / source callables used by the reference's tests and demos
(tests/test_solver.py:13-32, examples/*.py upstream) plus the extra ones of make_golden.py."""
import numpy as np


def top_hot(i, j, rows, cols):
    if i == 0:
        return 1.0
    if i == rows - 1 or j == 0 or j == cols - 1:
        return 0.0
    return None


def zero_boundary(i, j, rows, cols):
    if i == 0 or i == rows - 1 or j == 0 or j == cols - 1:
        return 0.0
    return None


def sin_source(i, j, rows, cols):
    x = i / (rows - 1)
    y = j / (cols - 1)
    return -2.0 * np.pi ** 2 * np.sin(np.pi * x) * np.sin(np.pi * y)


def warm_interior(i, j, rows, cols):
    if i == 0:
        return 1.0
    if i == rows - 1 or j == 0 or j == cols - 1:
        return 0.25
    return 0.5 + 0.01 * ((i * 7 + j * 3) % 11)


def partial_ring(i, j, rows, cols):
    if i == 0:
        return 1.0
    if j == 0:
        return 0.5
    return None


BCS = {"top_hot": top_hot, "zero_boundary": zero_boundary, "warm_interior": warm_interior,
       "partial_ring": partial_ring}
SRCS = {"none": None, "sin_source": sin_source}


def random_problem(rows, cols, sigma, seed, with_source=True):
    """Seeded synthetic inputs as arrays (what the C-ABI consumes)."""
    rng = np.random.default_rng(seed)
    grid0 = np.zeros((rows, cols))
    grid0[0, :] = 1.0
    grid0[1:-1, 1:-1] = 0.1 * rng.standard_normal((max(rows - 2, 0), max(cols - 2, 0)))
    nh = nv = None
    if sigma > 0:
        nh = 1.0 + sigma * rng.standard_normal((rows, cols - 1))
        nv = 1.0 + sigma * rng.standard_normal((rows - 1, cols))
    src = 1e-3 * rng.standard_normal((rows, cols)) if with_source else None
    return grid0, nh, nv, src
