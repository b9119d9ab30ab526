import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/_build/liboracle.so), built on demand.  Test infrastructure only."""
    from oracle import oracle_py as O
    O.build(ref=True)
    return O


@pytest.fixture(scope="session")
def golden_search():
    return np.load(os.path.join(GOLDEN, "ref_search_59sats.npz"))


@pytest.fixture(scope="session")
def golden_stages():
    return np.load(os.path.join(GOLDEN, "ref_stages.npz"))


def _have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_required():
    # -m gpu tests must FAIL (not skip) when the CUDA path is unavailable on a GPU box;
    # on the CPU-only container they are simply not selected (-m "not gpu").
    assert _have_cuda(), "CUDA device required for -m gpu tests"
    return True
