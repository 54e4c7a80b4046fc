"""ctypes binding of tests/hostemu/libsvo_hostemu.so (TEST INFRASTRUCTURE): the DEVICE source of the trace path
(svo_trace.cuh) compiled for the host by g++ through cuda_host_shim.h.  See emu.cpp."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.normpath(os.path.join(_HERE, "..", "..", "svo_raytracer_b200", "csrc"))
_LIB_PATH = os.path.join(_HERE, "libsvo_hostemu.so")
_CPP = [os.path.join(_HERE, f) for f in ("emu.cpp", "kernels_emu.cpp", "wavefront_emu.cpp", "gpu_build_emu.cpp", "transcode_emu.cpp", "simt_emu.cpp")] + [os.path.join(_CSRC, "svo_transcode.cpp")]
_SRCS = _CPP + [os.path.join(_HERE, f) for f in ("cuda_host_shim.h", "simt_emu.h", "emu_scene.h")] + \
    [os.path.join(_CSRC, f) for f in ("svo_trace.cuh", "detmath.cuh", "svo_kernels.h", "svo_kernels.cu", "svo_wavefront.cu", "svo_transcode.h", "svo_gpu_build.cu", "svo_gpu_build.h", "svo_gpu_transcode.cu", "svo_dev.h")]
CUDA_INCLUDE = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


def build(force: bool = False) -> str:
    stale = force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in _SRCS)
    if stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wno-unknown-pragmas", "-Wno-maybe-uninitialized", "-Wno-uninitialized",
                               "-I" + CUDA_INCLUDE, "-o", _LIB_PATH] + _CPP + ["-lpthread"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emu_scene_create.restype = C.c_void_p
        L.emu_scene_create.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_int]
        L.emu_scene_destroy.argtypes = [C.c_void_p]
        L.emu_scene_ndesc.restype = C.c_uint64
        L.emu_scene_ndesc.argtypes = [C.c_void_p]
        L.emu_render.restype = C.c_int
        L.emu_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_launch_render.restype = C.c_int
        L.emu_launch_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int]
        L.emu_gpu_build_terrain.restype = C.c_int
        L.emu_gpu_build_terrain.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
        L.emu_beam_conservative.restype = C.c_int
        L.emu_beam_conservative.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.emu_beam_in_parts.restype = C.c_int
        L.emu_beam_in_parts.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.emu_gpu_transcode_check.restype = C.c_int
        L.emu_gpu_transcode_check.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.emu_patch_check.restype = C.c_int
        L.emu_patch_check.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.emu_fence_word.restype = C.c_uint
        L.emu_fence_word.argtypes = [C.c_int]
        L.emu_launch_cast.restype = C.c_int
        L.emu_launch_cast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.emu_cast.restype = C.c_int
        L.emu_cast.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
        L.emu_beam.restype = C.c_int
        L.emu_beam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.emu_simt.restype = C.c_int
        L.emu_simt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.emu_pixel_ops.restype = C.c_int
        L.emu_pixel_ops.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.emu_simt_stream.restype = C.c_int
        L.emu_simt_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_selftest.restype = C.c_int
        L.emu_selftest.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.emu_math.restype = C.c_float
        L.emu_math.argtypes = [C.c_int, C.c_float, C.c_float]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


PATH_RUN, PATH_STEP, PATH_STATS = 0, 1, 2


class Scene:
    """A node stream transcoded by the product's host transcoder, viewed through the device structs."""

    def __init__(self, nodes: np.ndarray):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
        err = C.create_string_buffer(256)
        self._h = lib().emu_scene_create(_ptr(nodes), nodes.size, err, 256)
        if not self._h:
            raise ValueError(err.value.decode())

    def close(self):
        if self._h:
            lib().emu_scene_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def ndesc(self):
        return int(lib().emu_scene_ndesc(self._h))

    def render(self, frame, width, height, y0=0, y1=None, path=PATH_RUN, box=False, aux=True, beam=None, prev_rgba8=None,
               nthreads=8):
        """`frame`: any ctypes struct laid out like svo_frame (oracle.Frame or the product's).  Returns the planes
        (and the counters of k_render_stats for path=PATH_STATS)."""
        y1 = height if y1 is None else y1
        out = {"rgba8": np.zeros((height, width, 4), np.uint8) if prev_rgba8 is None else np.ascontiguousarray(prev_rgba8, np.uint8).copy(),
               "depth": np.zeros((height, width), np.float32)}
        want_aux = aux or path == PATH_STATS
        if want_aux:
            out["radiance"] = np.zeros((height, width, 4), np.float32)
            out["hit_id"] = np.zeros((height, width), np.uint32)
            out["iter"] = np.zeros((height, width), np.uint32)
            out["primary_t"] = np.zeros((height, width), np.float32)
        if beam is not None:
            beam = np.ascontiguousarray(beam, dtype=np.float32)
        counters = np.zeros(3, np.uint64)
        rc = lib().emu_render(self._h, C.byref(frame), width, height, y0, y1, path, int(box), int(aux), _ptr(beam),
                              _ptr(out["rgba8"]), _ptr(out["depth"]), _ptr(out.get("hit_id")), _ptr(out.get("iter")),
                              _ptr(out.get("primary_t")), _ptr(out.get("radiance")), _ptr(counters), nthreads)
        assert rc == 0
        if path == PATH_STATS:
            return out, {"casts": int(counters[0]), "iters": int(counters[1]), "record_bytes": int(counters[2])}
        return out

    def launch_render(self, frame, width, height, y0=0, y1=None, kernel=0, box=False, aux=True, beam=None, prev_rgba8=None,
                      band_stride=0, band_offset=0, band_rows=8, ctas=6, nthreads=8, into=None):
        """The product's launch_render() -- variant dispatch, grid arithmetic, the __global__ kernels with their warp- and
        block-level collectives -- on the coroutine SIMT emulator (kernels_emu.cpp).  `into`: planes to draw into."""
        y1 = height if y1 is None else y1
        if into is not None:
            out = into
        else:
            out = {"rgba8": np.zeros((height, width, 4), np.uint8) if prev_rgba8 is None else np.ascontiguousarray(prev_rgba8, np.uint8).copy(),
                   "depth": np.zeros((height, width), np.float32), "radiance": np.zeros((height, width, 4), np.float32),
                   "hit_id": np.zeros((height, width), np.uint32), "iter": np.zeros((height, width), np.uint32),
                   "primary_t": np.zeros((height, width), np.float32)}
        if beam is not None:
            beam = np.ascontiguousarray(beam, dtype=np.float32)
        rc = lib().emu_launch_render(self._h, C.byref(frame), width, height, y0, y1, kernel, int(box), int(aux), _ptr(beam),
                                     _ptr(out["rgba8"]), _ptr(out["depth"]), _ptr(out["hit_id"]), _ptr(out["iter"]),
                                     _ptr(out["primary_t"]), _ptr(out["radiance"]), band_stride, band_offset, band_rows, ctas, nthreads)
        assert rc == 0, rc
        self.last_launches = int(lib().emu_last_render_launches())
        return out

    def launch_cast(self, rays, max_depth=13, kernel=0, order=None, ctas=2, nthreads=8):
        """The product's launch_cast() (ray-stream kernels) on the SIMT emulator."""
        from oracle import oracle as O
        rays = np.ascontiguousarray(rays, dtype=O.RAY_DTYPE)
        out = np.zeros(rays.shape[0], dtype=O.HIT_DTYPE)
        if order is not None:
            order = np.ascontiguousarray(order, dtype=np.uint32)
        rc = lib().emu_launch_cast(self._h, _ptr(rays), _ptr(order), rays.shape[0], _ptr(out), max_depth, kernel, ctas, nthreads)
        assert rc == 0, rc
        return out

    def beam_conservative(self, frame, width, height, nthreads=8):
        """svo_beam_conservative on the emulator: per 4x4 block a lower bound on the primary hit distance (+inf = all miss)."""
        out = np.zeros((height // 4, width // 4), np.float32)
        rc = lib().emu_beam_conservative(self._h, C.byref(frame), width, height, _ptr(out), nthreads)
        assert rc == 0, rc
        return out

    def beam_in_parts(self, frame, width, height, parts, nthreads=8):
        """svo_beam_lattice_rows per part (each storing into two lattices and bumping two fences) + svo_beam_filter."""
        out = np.zeros((height // 4, width // 4), np.float32)
        rc = lib().emu_beam_in_parts(self._h, C.byref(frame), width, height, parts, _ptr(out), nthreads)
        assert rc == 0, rc
        return out

    def cast(self, rays, max_depth=13, nthreads=8):
        from oracle import oracle as O
        rays = np.ascontiguousarray(rays, dtype=O.RAY_DTYPE)
        out = np.zeros(rays.shape[0], dtype=O.HIT_DTYPE)
        lib().emu_cast(self._h, _ptr(rays), rays.shape[0], _ptr(out), max_depth, nthreads)
        return out

    def simt(self, frame, width, height, costs, y0=0, y1=None, box=True, nthreads=8, tile_w=8):
        """The SIMT divergence model of emu.cpp: issue slots of the traversal under several loop organisations."""
        y1 = height if y1 is None else y1
        costs = np.ascontiguousarray(costs, dtype=np.float64)
        if costs.size == 8:
            costs = np.concatenate([costs, [110.0, 0.0]])  # refill event, extra instructions per warp iteration of the refill what-if
        assert costs.size == 10
        out = np.zeros(48, np.float64)
        lib().emu_simt(self._h, C.byref(frame), width, height, y0, y1, int(box), _ptr(costs), _ptr(out), nthreads, tile_w)
        names = ("if_if", "while_while", "ww_1_1", "ww_inf_1", "ww_1_inf", "ww_4_2")
        r = {n: float(out[i]) for i, n in enumerate(names)}
        r.update(ideal=float(out[8]), longest_lane=float(out[9]), casts=int(out[10]), iters=int(out[11]), pushes=int(out[12]),
                 advances=int(out[13]), pops=int(out[14]), warp_casts=int(out[15]), warp_iters=int(out[16]),
                 warp_iters_push=int(out[17]), warp_iters_adv=int(out[18]), warp_iters_pop=int(out[19]), early=int(out[47]))
        for g, G in enumerate((4, 16, 64)):
            for o, name in enumerate(("as_is", "by_length", "by_dy", "by_oct_dy", "by_oct")):
                r["regroup%d_%s" % (G, name)] = float(out[20 + g * 5 + o])
            for o, name in enumerate(("half_warps_idle8", "half_warps_idle16", "quarter_warps_idle8", "quarter_warps_idle16")):
                r["refill%d_%s" % (G, name)] = float(out[35 + g * 4 + o])
        return r

    def simt_stream(self, rays, costs, order=None, max_depth=13, nthreads=8):
        """The divergence model for a ray stream: issue slots of the grid-stride and of the persistent (lane refill) kernel."""
        from oracle import oracle as O
        rays = np.ascontiguousarray(rays, dtype=O.RAY_DTYPE)
        costs = np.ascontiguousarray(costs, dtype=np.float64)
        assert costs.size == 10
        if order is not None:
            order = np.ascontiguousarray(order, dtype=np.uint32)
        out = np.zeros(8, np.float64)
        lib().emu_simt_stream(self._h, _ptr(rays), _ptr(order), rays.shape[0], max_depth, _ptr(costs), _ptr(out), nthreads)
        return {"grid_stride": float(out[0]), "persistent": float(out[1]), "ideal": float(out[2]), "casts": int(out[3]), "iters": int(out[4]),
                "longest_lane": float(out[5])}

    def pixel_ops(self, frame, width, height, x, y, box=True):
        """Path string of pixel (x, y): op letter (P/A/Q/H/M, X = ended before the loop) + scale letter per iteration."""
        buf = C.create_string_buffer(1 << 16)
        lib().emu_pixel_ops(self._h, C.byref(frame), width, height, x, y, int(box), buf, len(buf))
        return buf.value.decode()

    def beam(self, frame, width, height):
        out = np.zeros((height // 4, width // 4), np.float32)
        lib().emu_beam(self._h, C.byref(frame), _ptr(out), width, height)
        return out


def selftest(blocks=3, os_threads=2):
    """Runs the emulator's self-test kernel; returns out[block, thread, 8]."""
    out = np.zeros((blocks, 128, 8), np.uint32)
    lib().emu_selftest(_ptr(out), blocks, os_threads)
    return out


def math(fn: int, x: float, y: float = 0.0) -> float:
    return float(lib().emu_math(fn, float(x), float(y)))


def gpu_build_terrain(height, mat, n, chunk, nthreads=8):
    """svo_gpu_build.cu (world generation on the device) on the SIMT emulator.  Returns the node stream."""
    height = np.ascontiguousarray(height, dtype=np.uint16)
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    nb = C.c_uint64(0)
    rc = lib().emu_gpu_build_terrain(_ptr(height), _ptr(mat), n, chunk, None, 0, C.byref(nb), nthreads)
    assert rc == 0, rc
    out = np.zeros(int(nb.value), np.uint8)
    rc = lib().emu_gpu_build_terrain(_ptr(height), _ptr(mat), n, chunk, _ptr(out), out.size, C.byref(nb), nthreads)
    assert rc == 0, rc
    return out


def gpu_transcode_check(nodes, nthreads=8) -> int:
    """svo_gpu_transcode.cu's whole transcode on the emulator against the host transcode; 0 = identical."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    return int(lib().emu_gpu_transcode_check(_ptr(nodes), nodes.size, nthreads))


def patch_check(old, new, ranges, nthreads=8) -> dict:
    """Incremental transcode (gpu_diff_apply + gpu_patch) of `ranges` = [(start, end), ...] turning stream `old` into `new`."""
    old = np.ascontiguousarray(old, dtype=np.uint8)
    new = np.ascontiguousarray(new, dtype=np.uint8)
    r = np.ascontiguousarray(np.array(ranges, dtype=np.uint64).reshape(-1))
    out = np.zeros(8, np.uint64)
    rc = lib().emu_patch_check(_ptr(old), old.size, _ptr(new), new.size, _ptr(r), len(ranges), _ptr(out), nthreads)
    assert rc == 0, rc
    return {"status": int(out[0]), "dirty": int(out[1]), "roots": int(out[2]), "appended": int(out[3]), "fell_back": int(out[4]),
            "reachable": int(out[5]), "stored": int(out[6])}
