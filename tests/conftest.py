import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Make sure the in-tree artefacts exist (no-ops when they are up to date): the product library
    (nvcc, sm_100a; cross-compiles without a GPU) and the CPU oracle (gcc)."""
    for d in (os.path.join(ROOT, "sparse_linear_algebra_b200", "csrc"), os.path.join(ROOT, "oracle")):
        try:
            subprocess.run(["make", "-C", d, "-j8", "-s"], check=True, stdout=subprocess.DEVNULL)
        except Exception as e:                      # the tests that need the artefact will say so themselves
            sys.stderr.write(f"[conftest] could not build {d}: {e}\n")


@pytest.fixture(scope="session")
def ora():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle

    oracle.build()
    oracle.lib()
    return oracle
