"""GPU: the reference's OWN test-suite and demo scripts, unmodified, against the B200 backend
(SURVEY section 4 plan (i)).  The files are staged by oracle/stage_reference.py
into the git-ignored oracle/_ref/ (they are reference sources: never committed); skipped when absent."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT, PKG_DIR

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref")
LOG_DIR = os.path.join(ROOT, "gpurun_out")


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = PKG_DIR  # `import analog_pde_solver` resolves to the B200 mirror package
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env.pop("APDE_MODE", None)   # default mode = exact (drop-in parity)
    return env


def _log(name, text):
    try:
        os.makedirs(LOG_DIR, exist_ok=True)
        with open(os.path.join(LOG_DIR, name), "w") as f:
            f.write(text)
    except OSError:
        pass


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "reference_tests", "test_solver.py")),
                    reason="reference tests not staged (oracle/stage_reference.py needs /root/reference)")
def test_reference_test_suite_passes_unmodified_on_the_backend():
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider",
                        os.path.join(REF, "reference_tests", "test_solver.py")],
                       capture_output=True, text=True, timeout=900, env=_env(), cwd=REF)
    _log("reference_suite.log", r.stdout + r.stderr)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert re.search(r"\b21 passed\b", r.stdout), r.stdout[-500:]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "reference_examples", "laplace_heat_demo.py")),
                    reason="reference demos not staged")
def test_reference_demos_print_the_reference_numbers():
    """BASELINE configs 1 and 2 run from the reference's own scripts: the printed iteration counts, errors
    and energies are the ones the reference prints on the CPU ('Demo outputs')."""
    out = {}
    for name in ("laplace_heat_demo.py", "poisson_demo.py"):
        r = subprocess.run([sys.executable, os.path.join(REF, "reference_examples", name)],
                           capture_output=True, text=True, timeout=600, env=_env(), cwd=REF)
        _log(f"reference_{name}.log", r.stdout + r.stderr)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[name] = r.stdout
    lap, poi = out["laplace_heat_demo.py"], out["poisson_demo.py"]
    for needle in ("336", "355", "28.026", "1209600.000", "43159.2", "0.00e+00", "3.26e-02"):
        assert needle in lap, (needle, lap[-1500:])
    for needle in ("458", "2.2625e-03", "24.730", "1648800.000", "66671.7"):
        assert needle in poi, (needle, poi[-1500:])
