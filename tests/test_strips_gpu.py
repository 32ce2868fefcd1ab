"""GPU (needs >= 2 B200s, skipped otherwise): the row-strip driver over NCCL/NVLink equals the single-GPU
run bit for bit -- one process per GPU under torch.distributed.run, tools/strips_check.py."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT
from analog_pde_solver import _backend

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        return _backend.load().apde_device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two sm_100 GPUs")
def test_nccl_strips_bitwise_equal_single_gpu():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n = min(_n_gpus(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "strips_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "STRIPS_CHECK PASS" in r.stdout


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two sm_100 GPUs (on one GPU the call is the single-GPU solve)")
@pytest.mark.parametrize("mode,dt", [("redblack", "f64"), ("redblack32", "f32")])
def test_single_process_multi_gpu_solve_equals_single_gpu(mode, dt):
    """apde_solve_redblack_multi (one process, CUDA peer copies) == the single-GPU solve bit for bit."""
    import numpy as np
    from helpers import random_problem
    rows, cols = 700, 333
    g0, nh, nv, s = random_problem(rows, cols, 0.01, seed=44)
    ref = g0.copy()
    rh, rn, _ = _backend.solve(mode, ref, nh, nv, s, 0.0, 21, check_every=8, temporal_depth=3)
    for n in (2, 3, 8):
        g = g0.copy()
        h, it, _ = _backend.solve(mode, g, nh, nv, s, 0.0, 21, check_every=8, temporal_depth=3, n_gpus=n)
        assert it == rn == 21
        assert g.tobytes() == ref.tobytes(), f"n_gpus={n}"
        assert h.tobytes() == rh.tobytes(), f"n_gpus={n}"
    # early stop decision identical too
    a, b = g0.copy(), g0.copy()
    ha, na, _ = _backend.solve(mode, a, None, None, s, 1e-4, 4000, check_every=16)
    hb, nb, _ = _backend.solve(mode, b, None, None, s, 1e-4, 4000, check_every=16, n_gpus=4)
    assert na == nb and a.tobytes() == b.tobytes() and ha.tobytes() == hb.tobytes()


def test_multi_gpu_call_on_one_gpu_is_the_single_gpu_solve():
    """n_gpus larger than the box has degrades to the devices present (documented in include/apde.h);
    on a one-GPU box that is the plain single-GPU solve."""
    import numpy as np
    from helpers import random_problem
    g0, nh, nv, s = random_problem(90, 70, 0.01, seed=4)
    a, b = g0.copy(), g0.copy()
    ha, na, _ = _backend.solve("redblack", a, nh, nv, s, 0.0, 9, check_every=4, temporal_depth=2)
    hb, nb, _ = _backend.solve("redblack", b, nh, nv, s, 0.0, 9, check_every=4, temporal_depth=2, n_gpus=64)
    assert na == nb == 9 and a.tobytes() == b.tobytes() and ha.tobytes() == hb.tobytes()
