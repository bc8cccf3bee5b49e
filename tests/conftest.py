import os
import sys

# The tile-sharded path parks streams in flag waits (cuStreamWaitValue32).  Streams beyond the number of hardware queues
# share one, and a parked stream would then hold up whatever is queued behind it: give every stream its own queue
# (must be set before the CUDA context exists; bench.py does the same).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def spano_lib():
    """libspano.so, built on demand (nvcc cross-compiles without a GPU)."""
    from simplepanorama_b200 import build as sb
    sb.build()
    from simplepanorama_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def ctx(spano_lib):
    from simplepanorama_b200 import api
    c = api.Context(0)
    yield c
    c.close()
