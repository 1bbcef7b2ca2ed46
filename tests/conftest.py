import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (HERE, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The product never builds itself implicitly; the test session does (nvcc cross-compiles
    sm_100a without a GPU) so that a fresh checkout can run `pytest` directly."""
    from learning_drones_b200 import build as ldbuild

    ldbuild.build()


# ---- measured parity figures, printed at the end of the session (also with -q) -----------------------------------
PARITY_REPORT = []


@pytest.fixture(scope="session")
def parity_report():
    """Tests append one line per measured GPU-vs-oracle / GPU-vs-reference comparison: the number of drones whose
    integer outcome differs, not just a bound that passed."""
    return PARITY_REPORT.append


def pytest_terminal_summary(terminalreporter):
    if PARITY_REPORT:
        terminalreporter.write_sep("-", "measured parity (drones with a differing integer outcome)")
        for line in PARITY_REPORT:
            terminalreporter.write_line(line)
