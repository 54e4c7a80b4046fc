"""ctypes binding of oracle/_ref/libsvo_ref_java.so -- the reference's own octree builder and SDF brush (Octree.java,
OctreeThread.java, Util.java, sdf/*.java) compiled for the CPU by oracle/build_ref_java.py (TEST INFRASTRUCTURE; same import
rule as oracle.py: tests/ and tools that make fixtures only).  Same call shapes as oracle.build_dense / build_terrain."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_ref_java
from .oracle import _ptr

_lib = None


def available() -> bool:
    try:
        return build_ref_java.build() is not None
    except Exception:
        return False


def lib():
    global _lib
    if _lib is None:
        path = build_ref_java.build()
        if path is None:
            raise RuntimeError("oracle/_ref/libsvo_ref_java.so: no /root/reference to build from and no prebuilt library")
        L = C.CDLL(path)
        L.svo_refj_about.restype = C.c_char_p
        L.svo_refj_build_dense.restype = C.c_uint64
        L.svo_refj_build_dense.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.svo_refj_build_terrain.restype = C.c_uint64
        L.svo_refj_build_terrain.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.svo_refj_sdf_brush.restype = C.c_int
        L.svo_refj_sdf_brush.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p]
        _lib = L
    return _lib


def build_dense(voxels: np.ndarray, cap: int | None = None):
    """voxels[z, y, x] uint8 cube -> node bytes of one OctreeThread over the cube (constructInnerOctree as written)."""
    voxels = np.ascontiguousarray(voxels, dtype=np.uint8)
    n = voxels.shape[0]
    assert voxels.shape == (n, n, n)
    cap = cap or max(1 << 16, 16 * n * n * n)
    buf = np.zeros(cap, np.uint8)
    counts = (C.c_uint64 * 4)()
    used = lib().svo_refj_build_dense(_ptr(voxels), n, _ptr(buf), cap, C.byref(counts))
    if used == 0:
        raise MemoryError("reference builder: capacity too small")
    return buf[:used].copy(), [int(c) for c in counts]


def build_terrain(height: np.ndarray, mat: np.ndarray, n: int, chunk: int = 1024, cap: int | None = None):
    """constructCompleteOctree as written (fill levels, eight OctreeThreads per chunk, splice)."""
    height = np.ascontiguousarray(height, dtype=np.uint16)
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    assert height.shape == (n, n) and mat.shape == (n, n)
    cap = cap or max(1 << 20, 40 * n * n)
    buf = np.zeros(cap, np.uint8)
    counts = (C.c_uint64 * 4)()
    used = lib().svo_refj_build_terrain(_ptr(height), _ptr(mat), n, chunk, _ptr(buf), cap, C.byref(counts))
    if used == 0:
        raise MemoryError("reference builder: capacity %d too small" % cap)
    return buf[:used].copy(), [int(c) for c in counts]


def sdf_brush(nodes: np.ndarray, world_size: int, max_lod: int, origin, params, value: int, kind: str = "sphere", slack: int = 1 << 22):
    """Main.placeSDF: Octree.useSDFBrush(new Sphere(origin, r) | new Box(origin, w, h, d), value) on a copy of `nodes`.
    Returns (edited stream, [(start0, end0), (start1, end1)] with empty ranges dropped, raw ChangeBounds)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    buf = np.zeros(nodes.size + slack, np.uint8)
    buf[:nodes.size] = nodes
    nbytes = C.c_uint64(nodes.size)
    org = np.asarray(origin, np.int32)
    p = np.zeros(3, np.int32)
    p[:len(np.atleast_1d(params))] = np.atleast_1d(params)
    bounds = np.zeros(4, np.int64)
    rc = lib().svo_refj_sdf_brush(_ptr(buf), C.byref(nbytes), buf.size, world_size, max_lod, 0 if kind == "sphere" else 1, _ptr(org), _ptr(p),
                                  int(value), _ptr(bounds))
    if rc != 0:
        raise MemoryError("reference brush failed (%d): buffer too small" % rc)
    s0, e0, s1, e1 = (int(b) for b in bounds)
    ranges = [(a, b) for a, b in ((s0, e0), (s1, e1)) if b > a]
    return buf[:nbytes.value].copy(), ranges, (s0, e0, s1, e1)
