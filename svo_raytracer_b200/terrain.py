"""Deterministic synthetic terrain inputs (heightmap + material map).

The reference builds its worlds from a 16-bit heightmap PNG and an 8-bit
material map (``src/engine/Octree.java:206-226``; the 8192^2 maps it names are
absent upstream).  Benchmarks and tests use the procedural stand-in generated
by ``svo_terrain_generate`` (``csrc/svo_terrain.cpp``): same value ranges as
``assets/heightmaps/nz.png`` and ``assets/matmaps/materials.png``, integer
hashing plus IEEE double arithmetic, bit-reproducible for (n, seed).  Host-side
code; needs the built library but no GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

__all__ = ["terrain_inputs"]


def terrain_inputs(n: int, seed: int = 1, nthreads: int = 0):
    """(height u16 [n,n] with row = z, material u8 [n,n]) for an n^3 world."""
    height = np.empty((n, n), dtype=np.uint16)
    mat = np.empty((n, n), dtype=np.uint8)
    rc = L.lib().svo_terrain_generate(n, seed, height.ctypes.data_as(C.c_void_p), mat.ctypes.data_as(C.c_void_p), nthreads)
    if rc != L.OK:
        raise L.SvoError(rc, "svo_terrain_generate(n=%d) failed" % n)
    return height, mat
