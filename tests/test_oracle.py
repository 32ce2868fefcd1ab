"""CPU: the oracle (oracle/apde_oracle.c) against the golden vectors of the unmodified reference."""
import hashlib

import numpy as np
import pytest

import oracle
from conftest import GOLDEN_NAMES


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_bit_exact_vs_reference(golden, name):
    c = golden.case(name)
    u, hist, n = oracle.gauss_seidel(c["grid0"], c["noise_h"], c["noise_v"], c["source"], c["tol"], c["max_iter"])
    assert n == c["iters"]
    assert u.tobytes() == c["solution"].tobytes()
    assert hist.tobytes() == c["hist"].tobytes()
    assert hashlib.sha256(u.tobytes()).hexdigest()[:16] == c["meta"]["sha256_16"]


def test_survey_table_fingerprints(golden):
    """SURVEY.md section 4 golden table (iterations + sha of the solution bytes)."""
    expect = {"laplace20_ideal": (336, "7fc5f23945ff2141"), "laplace20_noisy": (355, "25acc3389334da80"),
              "laplace20_digital": (336, "7fc5f23945ff2141"), "poisson20_analog": (458, "1019e73d85f8f9ec"),
              "poisson64_digital": (4070, "046fed5aa6c56522"), "poisson64_noisy": (5718, "dc968241c02c6333"),
              "laplace10_tol5": (68, "2a4d91b635720580"), "laplace10_tol6": (87, "debc99457ebe5068"),
              "laplace10_noisy7": (90, "a0d3deb361902c5e"), "poisson15": (260, "73112524b4c7e6fe"),
              "laplace15_tol7": (239, "dd1ac4778ea6a3b7"), "rect12x30_noisy_src": (222, "a26b988a86d170ff")}
    for name, (iters, sha) in expect.items():
        assert golden.index[name]["iters"] == iters
        assert golden.index[name]["sha256_16"] == sha


def test_unit_links_equal_no_links(golden):
    """noise_level = 0 analog == digital bit for bit (x / 1.0 == x), SURVEY finding 3."""
    c = golden.case("poisson20_digital")
    ones_h = np.ones((20, 19))
    ones_v = np.ones((19, 20))
    a = oracle.gauss_seidel(c["grid0"], ones_h, ones_v, c["source"], c["tol"], c["max_iter"])
    b = oracle.gauss_seidel(c["grid0"], None, None, c["source"], c["tol"], c["max_iter"])
    assert a[2] == b[2] and a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()


@pytest.mark.parametrize("name", ["rect12x30_noisy_src", "mid96x80_k12", "mid150_k6_digital", "three_by_three", "rows2"])
@pytest.mark.parametrize("threads", [2, 3, 8])
def test_multithreaded_oracle_is_bitwise_serial(golden, name, threads):
    c = golden.case(name)
    k = min(c["iters"], 9)
    a = oracle.gauss_seidel(c["grid0"], c["noise_h"], c["noise_v"], c["source"], 0.0, k)
    b = oracle.gauss_seidel_mt(c["grid0"], c["noise_h"], c["noise_v"], c["source"], k, threads)
    assert a[0].tobytes() == b[0].tobytes()
    assert a[1].tobytes() == b[1].tobytes()


def test_red_black_same_fixed_point_as_reference_order():
    """tol ~ 1e-13: red-black and lexicographic agree to <= 1e-10 relative (north_star bar)."""
    from helpers import random_problem
    g0, nh, nv, src = random_problem(20, 20, 0.01, 3)
    gs, _, n = oracle.gauss_seidel(g0, nh, nv, src, 1e-14, 20000)
    assert n < 20000
    rb, hist, _ = oracle.red_black(g0, nh, nv, src, 2500)
    assert hist[-1] < 1e-13
    assert np.max(np.abs(rb - gs)) <= 1e-10 * np.max(np.abs(gs))


def test_red_black_no_noise_matches_numpy_statement():
    """Independent numpy statement of one red-black sweep (no mismatch): same bits."""
    rng = np.random.default_rng(0)
    u = rng.standard_normal((9, 14))
    src = rng.standard_normal((9, 14))
    ref = u.copy()
    for colour in (0, 1):
        for i in range(1, 8):
            for j in range(1, 13):
                if (i + j) % 2 == colour:
                    # fma(uS, 0.25, tq): 0.25*uS is exact, so the sum below rounds exactly once
                    tq = (((ref[i - 1, j] + ref[i, j - 1]) + ref[i, j + 1]) - src[i, j]) * 0.25
                    ref[i, j] = 0.25 * ref[i + 1, j] + tq
    got, hist, _ = oracle.red_black(u, None, None, src, 1)
    assert got.tobytes() == ref.tobytes()
    assert hist[0] == np.max(np.abs(ref - u))


def test_red_black_f32_close_to_f64():
    from helpers import random_problem
    g0, nh, nv, src = random_problem(64, 48, 0.01, 5)
    a, _, _ = oracle.red_black(g0, nh, nv, src, 50, dtype=np.float64)
    b, _, _ = oracle.red_black(g0, nh, nv, src, 50, dtype=np.float32)
    assert b.dtype == np.float32
    assert np.max(np.abs(a - b)) < 1e-5
