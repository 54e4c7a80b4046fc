"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- see svo_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsvo_oracle.so")
NO_HIT = 0xFFFFFFFF


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("svo_oracle.c", "svo_builder.c", "svo_oracle.h", "oracle_math.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsvo_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class CastResult(C.Structure):
    _fields_ = [("value", C.c_uint32), ("pointer", C.c_uint32), ("iter", C.c_uint32), ("t", C.c_float),
                ("hitPos", C.c_float * 3), ("scale", C.c_float), ("debugColor", C.c_float * 3),
                ("normal", C.c_float * 3), ("voxelPos", C.c_float * 3), ("depth", C.c_uint32)]


class Frame(C.Structure):
    _fields_ = [("camPos", C.c_float * 3), ("l1", C.c_float * 3), ("l2", C.c_float * 3),
                ("r1", C.c_float * 3), ("r2", C.c_float * 3),
                ("frameNumber", C.c_int32), ("renderMode", C.c_int32), ("useBeam", C.c_int32),
                ("maxDepth", C.c_int32), ("casts", C.c_int32), ("coneDepth", C.c_int32),
                ("mirrorValue", C.c_int32), ("flags", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("casts", C.c_uint64), ("iters", C.c_uint64), ("record_bytes", C.c_uint64),
                ("stale_pops", C.c_uint64), ("capped", C.c_uint64), ("hits", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3)])
HIT_DTYPE = np.dtype([("id", np.uint32), ("t", np.float32), ("value", np.uint32), ("iter", np.uint32)])

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.svo_oracle_cast.restype = C.c_int
        L.svo_oracle_cast.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.c_int, C.c_int, C.c_int, C.POINTER(CastResult), C.POINTER(Stats)]
        L.svo_oracle_cast_rays.restype = None
        L.svo_oracle_cast_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int,
                                           C.c_void_p, C.c_int, C.POINTER(Stats)]
        L.svo_oracle_render.restype = None
        L.svo_oracle_render.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Frame), C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Stats)]
        L.svo_oracle_beam.restype = None
        L.svo_oracle_beam.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Frame), C.c_int, C.c_int, C.c_void_p, C.c_int]
        for name in ("sin", "cos", "acos", "exp"):
            f = getattr(L, "svo_oracle_" + name)
            f.restype = C.c_float
            f.argtypes = [C.c_float]
        L.svo_oracle_rand.restype = C.c_float
        L.svo_oracle_rand.argtypes = [C.c_float, C.c_float]
        L.svo_oracle_build_terrain.restype = C.c_uint64
        L.svo_oracle_build_terrain.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                               C.POINTER(C.c_uint64 * 4)]
        L.svo_oracle_build_dense.restype = C.c_uint64
        L.svo_oracle_build_dense.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64 * 4)]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_frame(cam_pos, l1, l2, r1, r2, frame_number=1, render_mode=0, use_beam=0, max_depth=13, casts=2,
               cone_depth=11, mirror_value=0, flags=0) -> Frame:
    f = Frame()
    f.camPos[:] = [float(v) for v in cam_pos]
    f.l1[:] = [float(v) for v in l1]
    f.l2[:] = [float(v) for v in l2]
    f.r1[:] = [float(v) for v in r1]
    f.r2[:] = [float(v) for v in r2]
    f.frameNumber, f.renderMode, f.useBeam = int(frame_number), int(render_mode), int(use_beam)
    f.maxDepth, f.casts, f.coneDepth, f.mirrorValue, f.flags = int(max_depth), int(casts), int(cone_depth), int(mirror_value), int(flags)
    return f


def cast(nodes: np.ndarray, o, d, max_depth=13, cone_trace=False, cone_depth=11, res: CastResult | None = None):
    """One intersectOctree call.  Returns (hit: bool, CastResult, Stats)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    res = res if res is not None else CastResult()
    st = Stats()
    oo = (C.c_float * 3)(*[float(v) for v in o])
    dd = (C.c_float * 3)(*[float(v) for v in d])
    hit = lib().svo_oracle_cast(_ptr(nodes), nodes.size, oo, dd, max_depth, int(cone_trace), cone_depth,
                                C.byref(res), C.byref(st))
    return bool(hit), res, st


def cast_rays(nodes: np.ndarray, rays: np.ndarray, max_depth=13, nthreads=1):
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
    st = Stats()
    lib().svo_oracle_cast_rays(_ptr(nodes), nodes.size, _ptr(rays), rays.shape[0], max_depth, _ptr(out),
                               nthreads, C.byref(st))
    return out, st


def render(nodes: np.ndarray, frame: Frame, width: int, height: int, y0=0, y1=None, beam=None, nthreads=1,
           planes=("rgba8", "depth", "radiance", "hit_id", "iter", "primary_t"), prev_rgba8=None):
    """svotrace.comp main() over rows [y0, y1).  Returns (dict of planes, Stats).  prev_rgba8: the framebuffer
    content before the frame (only read when frame.flags bit 0 -- progressive accumulation -- is set)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    y1 = height if y1 is None else y1
    out = {}
    if "rgba8" in planes:
        out["rgba8"] = np.zeros((height, width, 4), np.uint8) if prev_rgba8 is None else np.ascontiguousarray(prev_rgba8, np.uint8).copy()
    if "depth" in planes:
        out["depth"] = np.zeros((height, width), np.float32)
    if "radiance" in planes:
        out["radiance"] = np.zeros((height, width, 4), np.float32)
    if "hit_id" in planes:
        out["hit_id"] = np.zeros((height, width), np.uint32)
    if "iter" in planes:
        out["iter"] = np.zeros((height, width), np.uint32)
    if "primary_t" in planes:
        out["primary_t"] = np.zeros((height, width), np.float32)
    st = Stats()
    if beam is not None:
        beam = np.ascontiguousarray(beam, dtype=np.float32)
    lib().svo_oracle_render(_ptr(nodes), nodes.size, C.byref(frame), width, height, y0, y1, _ptr(beam),
                            _ptr(out.get("rgba8")), _ptr(out.get("depth")), _ptr(out.get("radiance")),
                            _ptr(out.get("hit_id")), _ptr(out.get("iter")), _ptr(out.get("primary_t")),
                            nthreads, C.byref(st))
    return out, st


def beam(nodes: np.ndarray, frame: Frame, width: int, height: int):
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    out = np.zeros((height // 4, width // 4), np.float32)
    lib().svo_oracle_beam(_ptr(nodes), nodes.size, C.byref(frame), width, height, _ptr(out), 1)
    return out


def build_terrain(height: np.ndarray, mat: np.ndarray, n: int, chunk: int = 1024, cap: int | None = None):
    """Brute-force reference builder.  Returns (node bytes, counts[4])."""
    height = np.ascontiguousarray(height, dtype=np.uint16)
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    assert height.shape == (n, n) and mat.shape == (n, n)
    cap = cap or max(1 << 20, 40 * n * n)
    buf = np.zeros(cap, np.uint8)
    counts = (C.c_uint64 * 4)()
    used = lib().svo_oracle_build_terrain(_ptr(height), _ptr(mat), n, chunk, _ptr(buf), cap, C.byref(counts))
    if used == 0:
        raise MemoryError("oracle builder: capacity %d too small" % cap)
    return buf[:used].copy(), [int(c) for c in counts]


def build_dense(voxels: np.ndarray, cap: int | None = None):
    """voxels[z, y, x] uint8 cube -> node bytes of one OctreeThread over the cube."""
    voxels = np.ascontiguousarray(voxels, dtype=np.uint8)
    n = voxels.shape[0]
    assert voxels.shape == (n, n, n)
    cap = cap or max(1 << 16, 16 * n * n * n)
    buf = np.zeros(cap, np.uint8)
    counts = (C.c_uint64 * 4)()
    used = lib().svo_oracle_build_dense(_ptr(voxels), n, _ptr(buf), cap, C.byref(counts))
    if used == 0:
        raise MemoryError("oracle builder: capacity too small")
    return buf[:used].copy(), [int(c) for c in counts]
