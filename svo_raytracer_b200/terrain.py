"""Deterministic synthetic terrain inputs (heightmap + material map).

The reference builds its worlds from a 16-bit heightmap PNG and an 8-bit
material map (``/root/reference/src/engine/Octree.java:206-226``; the 8192^2
maps it names are absent upstream, ``.MISSING_LARGE_BLOBS``).  Benchmarks and
tests here use a procedural stand-in with the same value ranges as
``assets/heightmaps/nz.png`` (u16, ~3 %..90 % of full scale) and
``assets/matmaps/materials.png`` (1 = stone, 2 = scree, 3 = grass).

Everything is integer hashing plus IEEE add/mul in float64, so the arrays are
bit-reproducible on any host; no RNG library state is involved.
"""
from __future__ import annotations

import numpy as np

__all__ = ["heightmap", "material_map", "terrain_inputs"]


def _hash2(ix: np.ndarray, iy: np.ndarray, salt: int) -> np.ndarray:
    """32-bit integer mix of lattice coordinates -> uniform [0, 1) float64."""
    with np.errstate(over="ignore"):
        h = (ix.astype(np.uint32) * np.uint32(0x9E3779B1)) ^ (iy.astype(np.uint32) * np.uint32(0x85EBCA77))
        h ^= np.uint32(salt & 0xFFFFFFFF)
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x2C1B3C6D)
        h ^= h >> np.uint32(12)
        h *= np.uint32(0x297A2D39)
        h ^= h >> np.uint32(15)
    return h.astype(np.float64) * (1.0 / 4294967296.0)


def _value_noise(size: int, cells: int, salt: int) -> np.ndarray:
    """Smooth value noise with `cells` lattice cells across a size x size map."""
    # sample position of texel centre in lattice units (exact dyadic rationals)
    t = (np.arange(size, dtype=np.float64) + 0.5) * (float(cells) / float(size))
    i0 = np.floor(t).astype(np.int64)
    f = t - i0
    w = f * f * (3.0 - 2.0 * f)  # smoothstep
    i1 = i0 + 1
    gx0, gy0 = np.meshgrid(i0, i0, indexing="xy")
    gx1, gy1 = np.meshgrid(i1, i1, indexing="xy")
    wx, wy = np.meshgrid(w, w, indexing="xy")
    v00 = _hash2(gx0, gy0, salt)
    v10 = _hash2(gx1, gy0, salt)
    v01 = _hash2(gx0, gy1, salt)
    v11 = _hash2(gx1, gy1, salt)
    top = v00 + (v10 - v00) * wx
    bot = v01 + (v11 - v01) * wx
    return top + (bot - top) * wy


def heightmap(size: int, seed: int = 1) -> np.ndarray:
    """size x size uint16 heightmap, row index = z (world), column = x."""
    acc = np.zeros((size, size), dtype=np.float64)
    amp, total, cells, octave = 1.0, 0.0, 4, 0
    while cells <= size // 2 and octave < 11:
        n = _value_noise(size, cells, seed * 7919 + octave * 104729)
        if octave >= 2:
            n = 1.0 - np.abs(2.0 * n - 1.0)  # ridged detail
        acc += amp * n
        total += amp
        amp *= 0.5
        cells *= 2
        octave += 1
    acc /= total
    lo, hi = float(acc.min()), float(acc.max())
    acc = (acc - lo) / (hi - lo)
    acc = acc * (0.35 + 0.65 * acc)  # flatten valleys, keep sharp peaks
    # same span as nz.png: 1957 .. 58795 of 65535
    out = np.floor(1957.0 + acc * (58795.0 - 1957.0)).astype(np.uint16)
    return np.ascontiguousarray(out)


def material_map(height: np.ndarray) -> np.ndarray:
    """uint8 material ids (1 stone, 2 scree, 3 grass) from height and slope."""
    h = height.astype(np.int64)
    size = h.shape[0]
    # slope in height units per texel, scaled so the classes are resolution independent
    gx = np.abs(np.roll(h, -1, axis=1) - np.roll(h, 1, axis=1))
    gz = np.abs(np.roll(h, -1, axis=0) - np.roll(h, 1, axis=0))
    slope = (gx + gz) * size // 1024
    m = np.full(h.shape, 3, dtype=np.uint8)
    m[(slope > 500) | (h > 30000)] = 2
    m[(slope > 1100) | (h > 44000)] = 1
    return np.ascontiguousarray(m)


def terrain_inputs(n: int, seed: int = 1):
    """(height u16 [n,n], material u8 [n,n]) for an n^3 world."""
    hm = heightmap(n, seed)
    return hm, material_map(hm)
