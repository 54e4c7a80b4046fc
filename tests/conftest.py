import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_devices():
    try:
        import ctypes as C
        from svo_raytracer_b200 import _lib
        n = C.c_int(0)
        return n.value if _lib.lib().svo_device_count(C.byref(n)) != 0 else n.value
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the gpu-marked tests instead of failing them (the product
    has no CPU path: svo_create returns SVO_ERR_NO_DEVICE).  `-m gpu` on such a box still fails loudly."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product path has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def svo():
    import svo_raytracer_b200 as s
    return s


def _terrain(svo, oracle, n, chunk):
    hm, mm = svo.terrain_inputs(n)
    nodes, _ = oracle.build_terrain(hm, mm, n, chunk)
    return nodes


@pytest.fixture(scope="session")
def terrain128(svo, oracle):
    """128^3 terrain, 4 chunks of 64 per axis... (chunk 64 -> one fill level), oracle-built."""
    return _terrain(svo, oracle, 128, 64)


@pytest.fixture(scope="session")
def terrain512(svo, oracle):
    """BASELINE.json configs[0]: 512^3 terrain (single chunk)."""
    return _terrain(svo, oracle, 512, 512)
