import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "analog-pde-solver-sim_b200")
for p in (PKG_DIR, ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_built():
    """Build libapde.so / the oracle on first use (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(PKG_DIR, "lib", "libapde.so")
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, "oracle", "libapde_oracle.so")):
        import __graft_entry__
        __graft_entry__.build()


_ensure_built()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        from analog_pde_solver import _backend
        return _backend.load().apde_device_count() > 0
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no sm_100 GPU visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """tests/golden/golden_cases.npz: inputs + outputs of the unmodified reference."""

    def __init__(self):
        here = os.path.join(ROOT, "tests", "golden")
        self.blob = np.load(os.path.join(here, "golden_cases.npz"))
        with open(os.path.join(here, "golden_index.json")) as f:
            self.index = json.load(f)["cases"]

    def names(self):
        return list(self.index)

    def get(self, name, key):
        k = f"{name}/{key}"
        return self.blob[k] if k in self.blob.files else None

    def case(self, name):
        g = lambda k: self.get(name, k)  # noqa: E731
        return dict(grid0=g("grid0"), noise_h=g("noise_h"), noise_v=g("noise_v"), source=g("source"),
                    tol=float(g("tol")), max_iter=int(g("max_iter")), solution=g("solution"),
                    hist=g("hist"), iters=int(g("iters")), energy=float(g("energy")), meta=self.index[name])


_GOLDEN = Golden()
GOLDEN_NAMES = _GOLDEN.names()


@pytest.fixture(scope="session")
def golden():
    return _GOLDEN
