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
