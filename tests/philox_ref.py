"""Host statement (numpy) of the device link-mismatch generator of apde_plan_fill
(csrc/apde_redblack.cu: philox4x32_10 / philox_normal / fill_kernel).  TEST INFRASTRUCTURE.

Counter-based: link id -> Philox4x32-10(counter = [id_lo, id_hi, 'APDE', 0], key = [seed_lo, seed_hi])
-> two 64-bit uniforms -> Box-Muller cosine branch.  Link ids: horizontal link (i,j)-(i,j+1) has
id = i*cols + j, vertical link (i,j)-(i+1,j) has id = rows*cols + i*cols + j.  The conductance stored on
the device is g = 1 / (1 + sigma * z), the reciprocal of the reference's mismatch factor
(solver.py:209-210 upstream draws the factors from numpy's PCG64 stream instead; the device generator
is for benchmark-scale grids where uploading host arrays is impractical)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2            # 32 x 32 -> 64 bit products (operands < 2^32)
        hi0, lo0, hi1, lo1 = p0 >> S32, p0 & MASK, p1 >> S32, p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def normal(link_id, seed):
    link_id = np.asarray(link_id, dtype=np.uint64)
    z = np.zeros_like(link_id)
    c0, c1, c2, c3 = philox4x32_10(link_id & MASK, link_id >> S32, z + np.uint64(0x41504445), z,
                                   np.uint64(seed) & MASK, np.uint64(seed) >> S32)
    u1 = (c0.astype(np.float64) * 4294967296.0 + c1.astype(np.float64) + 0.5) * (1.0 / 18446744073709551616.0)
    u2 = (c2.astype(np.float64) * 4294967296.0 + c3.astype(np.float64) + 0.5) * (1.0 / 18446744073709551616.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def conductances(rows, cols, sigma, seed, row_lo=0, row_hi=None):
    """(gh, gv) as the plan stores them, float64, shape (row_hi-row_lo, cols): gh[:, -1] = 0, gv[rows-1] = 0."""
    row_hi = rows if row_hi is None else row_hi
    ii, jj = np.meshgrid(np.arange(row_lo, row_hi, dtype=np.uint64), np.arange(cols, dtype=np.uint64), indexing="ij")
    idh = ii * np.uint64(cols) + jj
    idv = idh + np.uint64(rows) * np.uint64(cols)
    gh = 1.0 / (1.0 + sigma * normal(idh, seed))
    gv = 1.0 / (1.0 + sigma * normal(idv, seed))
    gh[:, cols - 1] = 0.0
    gv[ii == np.uint64(rows - 1)] = 0.0
    return gh, gv
