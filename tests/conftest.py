import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, never pass silently on a fallback
    pass


@pytest.fixture(scope="session")
def orc():
    return graft.load_oracle()


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def capi(pkg):
    return pkg.capi


@pytest.fixture(scope="session")
def ctx(capi):
    """A GPU context; creation raises (test error) when the CUDA library or device is missing."""
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def c1():
    import workloads
    return workloads.load_c1()


@pytest.fixture(scope="session")
def bunny4k():
    import workloads
    return workloads.bunny_problem("moderate", seed=3, n_points=4167)


def rot_err(A, B):
    return float(np.arccos(np.clip((np.trace(A[:3, :3].T @ B[:3, :3]) - 1) / 2, -1, 1)))
