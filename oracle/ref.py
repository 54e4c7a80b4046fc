"""ctypes binding of oracle/_ref/libsvo_ref.so -- the reference's own shaders compiled for the CPU by
oracle/build_ref.py (TEST INFRASTRUCTURE; same import rule as oracle.py: tests/, smoke() and bench.py's CPU legs only).

Same call shapes as oracle.py so that a test can run `oracle.X(...)` and `ref.X(...)` on the same inputs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_ref
from .oracle import HIT_DTYPE, RAY_DTYPE, CastResult, Frame, _ptr, make_frame  # noqa: F401

_lib = None


def available() -> bool:
    """True when the library can be built here (/root/reference present) or was shipped prebuilt."""
    try:
        return build_ref.build() is not None
    except Exception:
        return False


def lib():
    global _lib
    if _lib is None:
        path = build_ref.build()
        if path is None:
            raise RuntimeError("oracle/_ref/libsvo_ref.so: no /root/reference to build from and no prebuilt library")
        L = C.CDLL(path)
        L.svo_ref_about.restype = C.c_char_p
        L.svo_ref_cast.restype = C.c_int
        L.svo_ref_cast.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                                   C.c_int, C.c_int, C.POINTER(CastResult), C.POINTER(C.c_uint32)]
        L.svo_ref_cast_rays.restype = None
        L.svo_ref_cast_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
        L.svo_ref_render.restype = C.c_int
        L.svo_ref_render.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Frame), C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int]
        L.svo_ref_beam.restype = C.c_int
        L.svo_ref_beam.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Frame), C.c_int, C.c_int, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def cast(nodes: np.ndarray, o, d, max_depth=13, cone_trace=False, cone_depth=11, res: CastResult | None = None):
    """One intersectOctree call of svotrace.comp.  Returns (hit, CastResult, loop iterations)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    res = res if res is not None else CastResult()
    oo = (C.c_float * 3)(*[float(v) for v in o])
    dd = (C.c_float * 3)(*[float(v) for v in d])
    it = C.c_uint32(0)
    hit = lib().svo_ref_cast(_ptr(nodes), nodes.size, oo, dd, max_depth, int(cone_trace), cone_depth, C.byref(res),
                             C.byref(it))
    return bool(hit), res, int(it.value)


def cast_rays(nodes: np.ndarray, rays: np.ndarray, max_depth=13, nthreads=1):
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
    lib().svo_ref_cast_rays(_ptr(nodes), nodes.size, _ptr(rays), rays.shape[0], max_depth, _ptr(out), nthreads)
    return out


def render(nodes: np.ndarray, frame: Frame, width: int, height: int, y0=0, y1=None, beam=None, nthreads=1,
           planes=("rgba8", "depth", "radiance", "hit_id", "iter", "primary_t")):
    """svotrace.comp main() over rows [y0, y1).  Returns the dict of planes (same keys/dtypes as oracle.render)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    y1 = height if y1 is None else y1
    shapes = {"rgba8": ((height, width, 4), np.uint8), "depth": ((height, width), np.float32),
              "radiance": ((height, width, 4), np.float32), "hit_id": ((height, width), np.uint32),
              "iter": ((height, width), np.uint32), "primary_t": ((height, width), np.float32)}
    out = {k: np.zeros(*shapes[k]) for k in planes}
    if beam is not None:
        beam = np.ascontiguousarray(beam, dtype=np.float32)
    rc = lib().svo_ref_render(_ptr(nodes), nodes.size, C.byref(frame), width, height, y0, y1, _ptr(beam),
                              _ptr(out.get("rgba8")), _ptr(out.get("depth")), _ptr(out.get("radiance")),
                              _ptr(out.get("hit_id")), _ptr(out.get("iter")), _ptr(out.get("primary_t")), nthreads)
    if rc != 0:
        raise ValueError("the shipped shader has no mirror material / progressive accumulation (frame.mirrorValue, flags)")
    return out


def beam(nodes: np.ndarray, frame: Frame, width: int, height: int):
    """svobeam.comp main() over the (width/4) x (height/4) beam image."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    out = np.zeros((height // 4, width // 4), np.float32)
    lib().svo_ref_beam(_ptr(nodes), nodes.size, C.byref(frame), width, height, _ptr(out), 1)
    return out
