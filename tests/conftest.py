import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """One stb context on cuda:0 for the whole GPU session."""
    from semtools_b200 import capi
    c = capi.Context(0)
    yield c
    dev, host = c.ticket_check()      # K1's dynamic tile schedule stayed consistent over the whole session
    assert dev == host
    c.close()


def unit_rows(rng, n, d=256):
    """Rows ~ N(0,1) L2-normalised in f32 (SURVEY 8d synthetic corpus)."""
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)
