import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    """On a machine without a usable GPU (or without the built library) the gpu-marked tests are
    skipped instead of erroring in the session fixture."""
    try:
        from semtools_b200 import capi
        have = capi.device_count() > 0
    except Exception:                                     # noqa: BLE001 - library missing
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible / libsemtools_b200.so not built")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
