"""Host-side mirror of the reference renderer boundary, over the C ABI.

Two layers:

* :class:`SvoContext` -- thin, explicit wrapper of ``include/svo_b200.h``
  (one method per entry point, numpy in / numpy out).
* :class:`Renderer` -- the reference's ``src/engine/Renderer.java`` interface
  for this path, same method names and argument meaning (``getInstance``,
  ``addShader``, ``useProgram``, ``addSSBO(7, buf)``, ``updateSSBO(7, buf,
  start, end)``, ``setUniformInteger``, ``dispatchCompute``, ``printGLErrors``)
  plus the raw ``glUniform3fv`` / ``glUniform1i`` / ``glGetTexImage`` calls
  ``src/engine/Main.java`` makes next to it (:132-146, :259-285).  The Java
  engine would bind the same C ABI through JNI/Panama; no JVM
  exists in this image, so this Python mirror is what the tests drive.

The reference toolchain (Java) is absent here, hence Python; nothing in this
module computes anything -- it forwards to libsvo_b200.so or raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import Frame, SvoError

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3)])
HIT_DTYPE = np.dtype([("id", np.uint32), ("t", np.float32), ("value", np.uint32), ("iter", np.uint32)])

_PLANE_DTYPE = {
    L.PLANE_COLOR_RGBA8: (np.uint8, 4), L.PLANE_DEPTH: (np.float32, 1), L.PLANE_BEAM: (np.float32, 1),
    L.PLANE_HIT_ID: (np.uint32, 1), L.PLANE_ITER: (np.uint32, 1), L.PLANE_PRIMARY_T: (np.float32, 1),
    L.PLANE_RADIANCE: (np.float32, 4),
}


def make_frame(cam_pos, l1, l2, r1, r2, frame_number=1, render_mode=2, use_beam=0, max_depth=13, casts=2,
               cone_depth=11, mirror_value=0, flags=0) -> Frame:
    """The uniforms of Main.java:269-283 as one struct (defaults = the shader's #defines)."""
    f = Frame()
    f.camPos[:] = [float(v) for v in cam_pos]
    f.l1[:] = [float(v) for v in l1]
    f.l2[:] = [float(v) for v in l2]
    f.r1[:] = [float(v) for v in r1]
    f.r2[:] = [float(v) for v in r2]
    f.frameNumber, f.renderMode, f.useBeam = int(frame_number), int(render_mode), int(use_beam)
    f.maxDepth, f.casts, f.coneDepth, f.mirrorValue, f.flags = int(max_depth), int(casts), int(cone_depth), int(mirror_value), int(flags)
    return f


class SvoContext:
    """One svo_ctx: a device, an image size, an uploaded octree."""

    def __init__(self, width: int, height: int, device: int = 0):
        self._lib = L.lib()
        self._h = C.c_void_p()
        self.width, self.height, self.device = int(width), int(height), int(device)
        rc = self._lib.svo_create(C.byref(self._h), device, width, height)
        if rc != L.OK:
            raise SvoError(rc, (self._lib.svo_last_error(None) or b"").decode())

    # -- plumbing -------------------------------------------------------------
    def _check(self, rc: int):
        if rc != L.OK:
            raise SvoError(rc, (self._lib.svo_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._lib.svo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def set_option(self, option: int, value: int):
        self._check(self._lib.svo_set_option(self._h, option, int(value)))

    def get_option(self, option: int) -> int:
        v = C.c_int64()
        self._check(self._lib.svo_get_option(self._h, option, C.byref(v)))
        return int(v.value)

    def set_stream(self, cuda_stream: Optional[int]):
        self._check(self._lib.svo_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    # -- octree ---------------------------------------------------------------
    def upload(self, nodes: np.ndarray):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
        self._check(self._lib.svo_upload(self._h, nodes.ctypes.data_as(C.c_void_p), nodes.size))

    def upload_ptr(self, ptr: int, nbytes: int):
        self._check(self._lib.svo_upload(self._h, C.c_void_p(ptr), nbytes))

    def upload_range(self, nodes: np.ndarray, start: int, end: int):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
        if end > nodes.size:
            raise ValueError("end beyond the buffer")
        self._check(self._lib.svo_upload_range(self._h, nodes.ctypes.data_as(C.c_void_p), int(start), int(end)))

    def build_terrain_device(self, height: np.ndarray, mat: np.ndarray, n: int, chunk: int = 1024) -> int:
        """World generation on the device (svo_build_terrain_device); the result becomes this context's scene.  Returns the stream size."""
        height = np.ascontiguousarray(height, dtype=np.uint16)
        mat = np.ascontiguousarray(mat, dtype=np.uint8)
        if height.shape != (n, n) or mat.shape != (n, n):
            raise ValueError("height and mat must be (n, n)")
        nb = C.c_uint64()
        self._check(self._lib.svo_build_terrain_device(self._h, height.ctypes.data_as(C.c_void_p), mat.ctypes.data_as(C.c_void_p), n, chunk, C.byref(nb)))
        return int(nb.value)

    def download(self) -> np.ndarray:
        """The scene's node stream back on the host."""
        out = np.empty(self.scene_info()["stream_bytes"], dtype=np.uint8)
        self._check(self._lib.svo_download(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def upload_stats(self) -> dict:
        """How the last upload_range was absorbed (svo_upload_stats)."""
        o = (C.c_uint64 * 4)()
        self._check(self._lib.svo_upload_stats(self._h, C.byref(o)))
        return {"dirty": int(o[0]), "roots": int(o[1]), "appended": int(o[2]), "whole_transcode": bool(o[3])}

    def scene_canonical(self) -> dict:
        """Layout-independent fingerprint of the descriptor tree (svo_scene_canonical)."""
        o = (C.c_uint64 * 4)()
        self._check(self._lib.svo_scene_canonical(self._h, C.byref(o)))
        return {"reachable": int(o[0]), "hash": int(o[1]), "depth": int(o[2]), "stored": int(o[3])}

    def scene_probe(self):
        """Hash / counts / bounds of the descriptors held on the device (same 8 words as svo_transcode_probe)."""
        out = (C.c_uint64 * 8)()
        self._check(self._lib.svo_scene_probe(self._h, C.byref(out)))
        return [int(v) for v in out]

    def scene_info(self) -> dict:
        info = (C.c_uint64 * 4)()
        self._check(self._lib.svo_scene_info(self._h, C.byref(info)))
        return {"stream_bytes": int(info[0]), "descriptors": int(info[1]), "levels": int(info[2]), "device_bytes": int(info[3])}

    # -- dispatch -------------------------------------------------------------
    def render(self, frame: Frame, y0: Optional[int] = None, y1: Optional[int] = None):
        if y0 is None and y1 is None:
            self._check(self._lib.svo_render(self._h, C.byref(frame)))
        else:
            self._check(self._lib.svo_render_rows(self._h, C.byref(frame), int(y0 or 0), int(self.height if y1 is None else y1)))

    def render_interleaved(self, frame: Frame, part: int, parts: int):
        """Render 8-row bands part, part+parts, ... in one launch (multi-GPU image partition)."""
        self._check(self._lib.svo_render_interleaved(self._h, C.byref(frame), int(part), int(parts)))

    def render_interleaved_signal(self, frame: Frame, part: int, parts: int, fence_ptrs: Sequence[int] = (), slot: int = 0):
        """render_interleaved + fence_signal(fence_ptrs, slot) in one launch (empty list: this context's own fence)."""
        arr = (C.c_void_p * max(1, len(fence_ptrs)))(*[C.c_void_p(p) for p in fence_ptrs])
        self._check(self._lib.svo_render_interleaved_signal(self._h, C.byref(frame), int(part), int(parts), arr, len(fence_ptrs), int(slot)))

    def beam(self, frame: Frame):
        self._check(self._lib.svo_beam(self._h, C.byref(frame)))

    def beam_conservative(self, frame: Frame):
        """Fill the beam plane with proven per-block lower bounds on the primary hit distance (svo_beam_conservative)."""
        self._check(self._lib.svo_beam_conservative(self._h, C.byref(frame)))

    def beam_lattice_rows(self, frame: Frame, row0: int, row1: int, dst_ptrs: Sequence[int] = (), fence_ptrs: Sequence[int] = (), slot: int = 0):
        """Lattice rows [row0, row1) of the conservative beam pre-pass into the given lattice buffers (own + peers'), then the fences."""
        d = (C.c_void_p * max(1, len(dst_ptrs)))(*[C.c_void_p(p) for p in dst_ptrs])
        f = (C.c_void_p * max(1, len(fence_ptrs)))(*[C.c_void_p(p) for p in fence_ptrs])
        self._check(self._lib.svo_beam_lattice_rows(self._h, C.byref(frame), int(row0), int(row1), d, len(dst_ptrs), f, len(fence_ptrs), int(slot)))

    def beam_filter(self):
        """Current lane's lattice -> its beam plane (min filter + margin)."""
        self._check(self._lib.svo_beam_filter(self._h))

    def sync(self):
        self._check(self._lib.svo_sync(self._h))

    # -- readback -------------------------------------------------------------
    def plane_shape(self, plane: int):
        dt, ch = _PLANE_DTYPE[plane]
        h, w = (self.height // 4, self.width // 4) if plane == L.PLANE_BEAM else (self.height, self.width)
        return ((h, w, ch) if ch > 1 else (h, w)), dt

    def read_plane(self, plane: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        shape, dt = self.plane_shape(plane)
        if out is None:
            out = np.empty(shape, dtype=dt)
        self._check(self._lib.svo_read_plane(self._h, plane, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def read_plane_into(self, plane: int, ptr: int, nbytes: int, y0: int = 0, y1: Optional[int] = None):
        """Read rows [y0,y1) of a plane into caller memory (e.g. a pinned torch tensor)."""
        rows = self.height // 4 if plane == L.PLANE_BEAM else self.height
        self._check(self._lib.svo_read_plane_rows(self._h, plane, y0, rows if y1 is None else y1, C.c_void_p(ptr), nbytes))

    def read_color_rgba8(self):
        return self.read_plane(L.PLANE_COLOR_RGBA8)

    def read_depth(self):
        return self.read_plane(L.PLANE_DEPTH)

    def read_depth_at(self, x: int, y: int) -> float:
        v = C.c_float()
        self._check(self._lib.svo_read_depth_at(self._h, x, y, C.byref(v)))
        return float(v.value)

    def read_hit_id(self):
        return self.read_plane(L.PLANE_HIT_ID)

    def read_iter(self):
        return self.read_plane(L.PLANE_ITER)

    def read_primary_t(self):
        return self.read_plane(L.PLANE_PRIMARY_T)

    def read_radiance(self):
        return self.read_plane(L.PLANE_RADIANCE)

    def read_planes_async(self, rgba8_ptr: int, depth_ptr: int):
        """Enqueue the read-back of the current colour/depth set into (pinned) host memory; see svo_read_planes_async."""
        self._check(self._lib.svo_read_planes_async(self._h, C.c_void_p(rgba8_ptr or 0), C.c_void_p(depth_ptr or 0)))

    def read_interleaved_async(self, part: int, parts: int, rgba8_ptr: int, depth_ptr: int):
        """This rank's bands of the current set into full-frame host buffers (svo_read_interleaved_async)."""
        self._check(self._lib.svo_read_interleaved_async(self._h, int(part), int(parts), C.c_void_p(rgba8_ptr), C.c_void_p(depth_ptr)))

    def swap_buffers(self):
        self._check(self._lib.svo_swap_buffers(self._h))

    def select_lane(self, lane: int):
        """Make lane 0 / 1 (stream + colour/depth set) current; frames on different lanes may overlap on the GPU."""
        self._check(self._lib.svo_select_lane(self._h, int(lane)))

    def read_wait(self):
        self._check(self._lib.svo_read_wait(self._h))

    def device_ptr(self, plane: int) -> int:
        return int(self._lib.svo_device_ptr(self._h, plane) or 0)

    def bind_plane(self, plane: int, device_ptr: Optional[int]):
        self._check(self._lib.svo_bind_plane(self._h, plane, C.c_void_p(device_ptr or 0)))

    def ipc_export(self, plane: int) -> bytes:
        """72-byte IPC handle of one of this context's own planes (send it to the peer processes)."""
        buf = (C.c_uint8 * 72)()
        self._check(self._lib.svo_ipc_export(self._h, plane, buf))
        return bytes(buf)

    def ipc_import(self, handle: bytes) -> int:
        """Map a peer's plane; the returned device pointer can be given to bind_plane."""
        buf = (C.c_uint8 * 72).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self._lib.svo_ipc_import(self._h, buf, C.byref(p)))
        return int(p.value)

    def fence_export(self) -> bytes:
        buf = (C.c_uint8 * 72)()
        self._check(self._lib.svo_fence_export(self._h, buf))
        return bytes(buf)

    def fence_signal(self, fence_ptrs: Sequence[int] = (), slot: int = 0):
        """Enqueue: add 1 to slot `slot` of each given (imported) fence counter; none given = this context's own."""
        n = len(fence_ptrs)
        arr = (C.c_void_p * max(n, 1))(*[C.c_void_p(p) for p in fence_ptrs])
        self._check(self._lib.svo_fence_signal(self._h, arr, n, slot))

    def fence_wait(self, target: int, slot: int = 0):
        """Enqueue: wait until slot `slot` of this context's own counter reaches `target` (modulo 2^32)."""
        self._check(self._lib.svo_fence_wait(self._h, slot, target & 0xFFFFFFFF))

    def fence_wait_signal(self, target: int, slot: int, fence_ptrs: Sequence[int], signal_slot: int = 0):
        """fence_wait(target, slot) then fence_signal(fence_ptrs, signal_slot) in one launch."""
        arr = (C.c_void_p * len(fence_ptrs))(*[C.c_void_p(p) for p in fence_ptrs])
        self._check(self._lib.svo_fence_wait_signal(self._h, int(slot), target & 0xFFFFFFFF, arr, len(fence_ptrs), int(signal_slot)))

    def fence_device_ptr(self) -> int:
        return int(self._lib.svo_fence_device_ptr(self._h) or 0)

    def fence_reset(self):
        self._check(self._lib.svo_fence_reset(self._h))

    def ipc_close(self, device_ptr: int):
        self._check(self._lib.svo_ipc_close(self._h, C.c_void_p(device_ptr)))

    # -- ray streams ----------------------------------------------------------
    def cast(self, rays: np.ndarray, max_depth: int = 13) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        self._check(self._lib.svo_cast(self._h, rays.ctypes.data_as(C.c_void_p), rays.shape[0],
                                       out.ctypes.data_as(C.c_void_p), max_depth))
        return out

    def cast_device(self, d_rays: int, n: int, d_out: int, max_depth: int = 13):
        self._check(self._lib.svo_cast_device(self._h, C.c_void_p(d_rays), n, C.c_void_p(d_out), max_depth))

    # -- measurement ----------------------------------------------------------
    def timer_begin(self):
        self._check(self._lib.svo_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self._check(self._lib.svo_timer_end(self._h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.svo_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def render_stats(self, frame: Frame) -> dict:
        """Instrumented render: the oracle's counters for this frame (see svo_render_stats)."""
        cnt = (C.c_uint64 * 3)()
        self._check(self._lib.svo_render_stats(self._h, C.byref(frame), C.byref(cnt)))
        return {"casts": int(cnt[0]), "iters": int(cnt[1]), "record_bytes": int(cnt[2])}

    def render_stats_executed(self, frame: Frame) -> dict:
        """Counters of what the production kernel executes on this frame (content box on; see svo_render_stats_executed)."""
        cnt = (C.c_uint64 * 3)()
        self._check(self._lib.svo_render_stats_executed(self._h, C.byref(frame), C.byref(cnt)))
        return {"casts": int(cnt[0]), "iters": int(cnt[1]), "record_bytes": int(cnt[2])}

    def gather_probe(self, working_set_bytes: int, loads_per_thread: int = 256) -> float:
        """Random 32-byte-sector gather rate (sectors/s) over a working set: the gather roofline."""
        v = C.c_double()
        self._check(self._lib.svo_gather_probe(self._h, int(working_set_bytes), int(loads_per_thread), C.byref(v)))
        return float(v.value)

    def math_probe(self, fn: int, x: np.ndarray, y: Optional[np.ndarray] = None) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float32)
        out = np.empty_like(x)
        self._check(self._lib.svo_math_probe(self._h, fn, x.ctypes.data_as(C.c_void_p),
                                             None if y is None else y.ctypes.data_as(C.c_void_p),
                                             out.ctypes.data_as(C.c_void_p), x.size))
        return out


def build_terrain(height: np.ndarray, mat: np.ndarray, n: int, chunk: int = 1024, nthreads: int = 0) -> np.ndarray:
    """World generation (svo_build_terrain): heightmap + material map -> node stream bytes."""
    lib = L.lib()
    height = np.ascontiguousarray(height, dtype=np.uint16)
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    if height.shape != (n, n) or mat.shape != (n, n):
        raise ValueError("height and mat must be (n, n)")
    need = C.c_uint64()
    # the builder is deterministic: size first, then fill
    rc = lib.svo_build_terrain(height.ctypes.data_as(C.c_void_p), mat.ctypes.data_as(C.c_void_p), n, chunk, None, 0,
                               C.byref(need), nthreads)
    if rc != L.OK:
        raise SvoError(rc, "svo_build_terrain(size) failed")
    out = np.empty(int(need.value), dtype=np.uint8)
    rc = lib.svo_build_terrain(height.ctypes.data_as(C.c_void_p), mat.ctypes.data_as(C.c_void_p), n, chunk,
                               out.ctypes.data_as(C.c_void_p), out.size, C.byref(need), nthreads)
    if rc != L.OK:
        raise SvoError(rc, "svo_build_terrain failed")
    return out


# ---------------------------------------------------------------------------
# Renderer.java mirror
# ---------------------------------------------------------------------------
class Shader:
    """Renderer.Shader (Renderer.java:18-28): a named compute program."""

    def __init__(self, name: str, kind: str):
        self.name = name
        self.kind = kind  # "trace" | "beam"


class Renderer:
    """Drop-in for src/engine/Renderer.java on the trace path.

    Only what the hot path uses is provided; GL-only helpers (textures for the
    world-gen voxeliser, the present pass) are out of scope.
    """

    _instance: Optional["Renderer"] = None
    SSBO_OCTREE = 7  # layout(std430, binding = 7), svotrace.comp:13

    def __init__(self, width: int = 1920, height: int = 1080, device: int = 0):
        # Constants.WINDOW_WIDTH/HEIGHT (Constants.java:4-5)
        self.ctx = SvoContext(width, height, device)
        self.shaders: list[Shader] = []
        self.current: Optional[Shader] = None
        self._frame = make_frame((0, 0, 0), (-1, -1, -1), (-1, 1, -1), (1, -1, -1), (1, 1, -1), frame_number=0)
        self._errors: list[str] = []
        self._ssbo: Optional[np.ndarray] = None

    @classmethod
    def getInstance(cls, width: int = 1920, height: int = 1080, device: int = 0) -> "Renderer":
        if cls._instance is None:
            cls._instance = cls(width, height, device)
        return cls._instance

    @classmethod
    def resetInstance(cls):
        if cls._instance is not None:
            cls._instance.ctx.close()
        cls._instance = None

    # Renderer.java:43-54.  The GLSL source path selects the CUDA kernel that replaces it.
    def addShader(self, name: str, path: str) -> Shader:
        base = path.replace("\\", "/").rsplit("/", 1)[-1]
        if base == "svotrace.comp":
            kind = "trace"
        elif base == "svobeam.comp":
            kind = "beam"
        else:
            self._errors.append("addShader: %s has no sm_100a replacement" % path)
            kind = "none"
        sh = Shader(name, kind)
        self.shaders.append(sh)
        return sh

    def getShaderByName(self, name: str) -> Optional[Shader]:  # Renderer.java:152-158
        for s in self.shaders:
            if s.name == name:
                return s
        return None

    def useProgram(self, shader: Shader):  # Renderer.java:114-116
        self.current = shader

    # Renderer.java:123-129
    def addSSBO(self, bindIndex: int, data: np.ndarray, length: Optional[int] = None):
        if bindIndex != self.SSBO_OCTREE:
            self._errors.append("addSSBO: binding %d is not on the trace path" % bindIndex)
            return
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n = data.size if length is None else int(length)  # Octree.memOffset (uniform 9, Main.java:278)
        try:
            self.ctx.upload(data[:n])
            self._ssbo = data
        except SvoError as e:
            self._errors.append(str(e))

    # Renderer.java:131-146
    def updateSSBO(self, bindIndex: int, data: np.ndarray, start: Optional[int] = None, end: Optional[int] = None):
        if bindIndex != self.SSBO_OCTREE:
            self._errors.append("updateSSBO: binding %d is not on the trace path" % bindIndex)
            return
        data = np.ascontiguousarray(data, dtype=np.uint8)
        if start is None:
            return self.addSSBO(bindIndex, data)
        if start >= end:
            print("Update SSBO error: Invalid parameters.")  # Renderer.java:137-140
            return
        try:
            self.ctx.upload_range(data, start, end)
        except SvoError as e:
            self._errors.append(str(e))

    # raw GL calls of Main.java:259-283
    def glUniform3fv(self, location: int, v: Sequence[float]):
        f = self._frame
        target = {8: f.camPos, 1: f.l1, 2: f.l2, 3: f.r1, 4: f.r2}.get(location)
        if target is None:
            self._errors.append("glUniform3fv: no vec3 uniform at location %d" % location)
            return
        target[:] = [float(x) for x in v]

    def glUniform1i(self, location: int, value: int):
        f = self._frame
        if location == 5:
            f.frameNumber = int(value)
        elif location == 6:
            f.renderMode = int(value)
        elif location == 9:
            pass  # bufferEnd: only feeds an unused global upstream (svotrace.comp:17,20)
        elif location == 11:
            f.useBeam = int(value)
        else:
            self._errors.append("glUniform1i: no int uniform at location %d" % location)

    def setUniformInteger(self, location: int, value: int):  # Renderer.java:56-58
        self.glUniform1i(location, value)

    # Renderer.java:118-121.  Group counts are accepted for signature parity; the
    # image size given at construction decides the launch.
    def dispatchCompute(self, shader: Shader, numGroupsX: int, numGroupsY: int, numGroupsZ: int):
        try:
            if shader.kind == "trace":
                self.ctx.render(self._frame)
            elif shader.kind == "beam":
                self.ctx.beam(self._frame)
            else:
                self._errors.append("dispatchCompute: shader %s has no kernel" % shader.name)
        except SvoError as e:
            self._errors.append(str(e))

    # glGetTexImage(depthbuffer, GL_RED, GL_FLOAT) of Main.java:132-146
    def getDepthImage(self) -> np.ndarray:
        return self.ctx.read_depth()

    def getFramebufferImage(self) -> np.ndarray:
        return self.ctx.read_color_rgba8()

    def printGLErrors(self):  # Renderer.java:160-165
        for e in self._errors:
            print("SVO ERR: " + e)
        n = len(self._errors)
        self._errors.clear()
        return n
