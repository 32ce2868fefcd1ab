#!/usr/bin/env python3
"""Stage the reference's own test-suite and demo scripts for the GPU box.  TEST INFRASTRUCTURE.

The GPU box has no /root/reference, and reference sources must not enter the repository.  Like a compiled
`oracle/_ref` checker, the UNMODIFIED reference tests (tests/test_solver.py upstream, 21 property tests) and
the two demo scripts are therefore copied -- when the reference tree is present, i.e. in the build
container -- into `oracle/_ref/`, which is git-ignored (never committed) but travels with the `gpurun`
snapshot.  `tests/test_reference_suite_gpu.py` then runs them against the B200 backend (SURVEY section 4,
plan (i)) and skips when nothing was staged.  The product never reads this directory."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [("tests/test_solver.py", "reference_tests/test_solver.py"),
         ("tests/__init__.py", "reference_tests/__init__.py"),
         ("examples/laplace_heat_demo.py", "reference_examples/laplace_heat_demo.py"),
         ("examples/poisson_demo.py", "reference_examples/poisson_demo.py")]


def stage(reference_root: str = "/root/reference") -> bool:
    if not os.path.isdir(reference_root):
        return False
    for src, dst in FILES:
        s, d = os.path.join(reference_root, src), os.path.join(DEST, dst)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    return True


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("staged into oracle/_ref" if ok else "reference tree not present: nothing staged")
