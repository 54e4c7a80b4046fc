import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


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
