"""ctypes loader for libsvo_b200.so (the C ABI in include/svo_b200.h).

The product path has no CPU fallback: if the shared library is missing this
module raises, and every compute entry point of the library itself fails with
SVO_ERR_NO_DEVICE when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVO_B200_LIB") or os.path.join(_HERE, "libsvo_b200.so")  # override: A/B builds in development

NO_HIT = 0xFFFFFFFF

# enum svo_status
OK, ERR_INVALID, ERR_OOM, ERR_NO_DEVICE, ERR_CUDA, ERR_FORMAT, ERR_NO_SCENE, ERR_FENCE = 0, 1, 2, 100, 999, 1000, 1001, 1002
# enum svo_plane
PLANE_COLOR_RGBA8, PLANE_DEPTH, PLANE_BEAM, PLANE_HIT_ID, PLANE_ITER, PLANE_PRIMARY_T, PLANE_RADIANCE, PLANE_BEAM_LATTICE = range(8)
PLANE_BACK = 0x100
# enum svo_option
OPT_AUX_PLANES, OPT_FAST_MATH, OPT_KERNEL, OPT_L2_PERSIST, OPT_RAY_SORT, OPT_CONTENT_BOUNDS, OPT_BAND_ROWS, OPT_GPU_TRANSCODE = 1, 2, 3, 4, 5, 6, 7, 8
OPT_STREAM_KERNEL = 9


class SvoError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("svo_b200 error %d: %s" % (code, message))
        self.code = code


class Frame(C.Structure):
    """struct svo_frame (include/svo_b200.h)."""
    _fields_ = [("camPos", C.c_float * 3), ("l1", C.c_float * 3), ("l2", C.c_float * 3),
                ("r1", C.c_float * 3), ("r2", C.c_float * 3),
                ("frameNumber", C.c_int32), ("renderMode", C.c_int32), ("useBeam", C.c_int32),
                ("maxDepth", C.c_int32), ("casts", C.c_int32), ("coneDepth", C.c_int32),
                ("mirrorValue", C.c_int32), ("flags", C.c_int32)]


# every symbol include/svo_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _i = C.c_void_p, C.c_uint64, C.c_int
SYMBOLS = {
    "svo_abi_version": (_i, []),
    "svo_device_count": (_i, [C.POINTER(_i)]),
    "svo_create": (_i, [C.POINTER(_vp), _i, _i, _i]),
    "svo_destroy": (None, [_vp]),
    "svo_last_error": (C.c_char_p, [_vp]),
    "svo_set_option": (_i, [_vp, _i, C.c_int64]),
    "svo_get_option": (_i, [_vp, _i, C.POINTER(C.c_int64)]),
    "svo_set_stream": (_i, [_vp, _vp]),
    "svo_upload": (_i, [_vp, _vp, _u64]),
    "svo_upload_range": (_i, [_vp, _vp, _u64, _u64]),
    "svo_build_terrain_device": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(_u64)]),
    "svo_download": (_i, [_vp, _vp, _u64]),
    "svo_upload_stats": (_i, [_vp, C.POINTER(_u64 * 4)]),
    "svo_scene_canonical": (_i, [_vp, C.POINTER(_u64 * 4)]),
    "svo_scene_info": (_i, [_vp, C.POINTER(_u64 * 4)]),
    "svo_render": (_i, [_vp, C.POINTER(Frame)]),
    "svo_render_rows": (_i, [_vp, C.POINTER(Frame), _i, _i]),
    "svo_render_interleaved": (_i, [_vp, C.POINTER(Frame), _i, _i]),
    "svo_render_interleaved_signal": (_i, [_vp, C.POINTER(Frame), _i, _i, C.POINTER(_vp), _i, _i]),
    "svo_beam": (_i, [_vp, C.POINTER(Frame)]),
    "svo_beam_conservative": (_i, [_vp, C.POINTER(Frame)]),
    "svo_beam_lattice_rows": (_i, [_vp, C.POINTER(Frame), _i, _i, C.POINTER(_vp), _i, C.POINTER(_vp), _i, _i]),
    "svo_beam_filter": (_i, [_vp]),
    "svo_sync": (_i, [_vp]),
    "svo_read_plane": (_i, [_vp, _i, _vp, _u64]),
    "svo_read_plane_rows": (_i, [_vp, _i, _i, _i, _vp, _u64]),
    "svo_read_color_rgba8": (_i, [_vp, _vp]),
    "svo_read_depth": (_i, [_vp, _vp]),
    "svo_read_depth_at": (_i, [_vp, _i, _i, C.POINTER(C.c_float)]),
    "svo_read_hit_id": (_i, [_vp, _vp]),
    "svo_read_iter": (_i, [_vp, _vp]),
    "svo_read_primary_t": (_i, [_vp, _vp]),
    "svo_read_radiance_f32": (_i, [_vp, _vp]),
    "svo_read_planes_async": (_i, [_vp, _vp, _vp]),
    "svo_read_interleaved_async": (_i, [_vp, _i, _i, _vp, _vp]),
    "svo_swap_buffers": (_i, [_vp]),
    "svo_select_lane": (_i, [_vp, _i]),
    "svo_read_wait": (_i, [_vp]),
    "svo_device_ptr": (_vp, [_vp, _i]),
    "svo_bind_plane": (_i, [_vp, _i, _vp]),
    "svo_ipc_export": (_i, [_vp, _i, _vp]),
    "svo_ipc_import": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "svo_ipc_close": (_i, [_vp, _vp]),
    "svo_fence_export": (_i, [_vp, _vp]),
    "svo_fence_device_ptr": (_vp, [_vp]),
    "svo_fence_signal": (_i, [_vp, C.POINTER(_vp), _i, _i]),
    "svo_fence_wait": (_i, [_vp, _i, C.c_uint32]),
    "svo_fence_wait_signal": (_i, [_vp, _i, C.c_uint32, C.POINTER(_vp), _i, _i]),
    "svo_fence_reset": (_i, [_vp]),
    "svo_cast": (_i, [_vp, _vp, _u64, _vp, _i]),
    "svo_cast_device": (_i, [_vp, _vp, _u64, _vp, _i]),
    "svo_timer_begin": (_i, [_vp]),
    "svo_timer_end": (_i, [_vp, C.POINTER(C.c_float)]),
    "svo_launch_count": (_i, [_vp, C.POINTER(_u64)]),
    "svo_render_stats": (_i, [_vp, C.POINTER(Frame), C.POINTER(_u64 * 3)]),
    "svo_render_stats_executed": (_i, [_vp, C.POINTER(Frame), C.POINTER(_u64 * 3)]),
    "svo_gather_probe": (_i, [_vp, _u64, _i, C.POINTER(C.c_double)]),
    "svo_math_probe": (_i, [_vp, _i, _vp, _vp, _vp, _u64]),
    "svo_terrain_generate": (_i, [_i, _i, _vp, _vp, _i]),
    "svo_scene_probe": (_i, [_vp, C.POINTER(_u64 * 8)]),
    "svo_transcode_probe": (_i, [_vp, _u64, _i, C.POINTER(_u64 * 8), _vp, _u64]),
    "svo_build_terrain": (_i, [_vp, _vp, _i, _i, _vp, _u64, C.POINTER(_u64), _i]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA extension; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libsvo_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C svo_raytracer_b200/csrc`. There is no CPU fallback for this path." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
