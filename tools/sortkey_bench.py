#!/usr/bin/env python
"""SVO_OPT_RAY_SORT 0 / 1 / 2 on a dense-origin incoherent stream (BASELINE configs[3] in spirit: many random directions from
every surface point of a patch), numpy + the CUDA runtime through ctypes only.  Times svo_cast_device on device-resident buffers.   usage: python tools/sortkey_bench.py [size=2048] [patch=1024] [dirs=8]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import svo_raytracer_b200 as svo  # noqa: E402
from svo_raytracer_b200 import _lib as L  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
patch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dirs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
hm, mm = svo.terrain_inputs(size)
ctx = svo.SvoContext(64, 64)
ctx.build_terrain_device(hm, mm, size, min(size, 1024))
depth = min(13, int(np.log2(size)))
rng = np.random.default_rng(42)
g = np.arange(patch) + (size - patch) // 2
gx, gz = np.meshgrid(g, g)
xz = np.tile(np.stack([gx.ravel(), gz.ravel()], 1), (dirs, 1))
n = xz.shape[0]
h = ((hm.astype(np.uint32) * (size // 4)) >> 16).astype(np.float32)
rays = np.zeros(n, dtype=svo.RAY_DTYPE)
rays["o"][:, 0] = (xz[:, 0] + 0.5) / size + 1.0
rays["o"][:, 1] = (h[xz[:, 1], xz[:, 0]] + 2.5) / size + 1.0
rays["o"][:, 2] = (xz[:, 1] + 0.5) / size + 1.0
d = rng.normal(size=(n, 3)).astype(np.float32)
rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True)
perm = rng.permutation(n)
rays = rays[perm]  # the caller's order is arbitrary
# device-resident buffers through the CUDA runtime (ctypes; no torch): the timed region is binning + traversal only
import ctypes as C  # noqa: E402
rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
d_rays, d_hits = C.c_void_p(), C.c_void_p()
assert rt.cudaMalloc(C.byref(d_rays), rays.nbytes) == 0 and rt.cudaMalloc(C.byref(d_hits), n * 16) == 0
assert rt.cudaMemcpy(d_rays, rays.ctypes.data_as(C.c_void_p), rays.nbytes, 1) == 0
hits = np.zeros(n, dtype=svo.HIT_DTYPE)
ref = None
for sort in (0, 1, 2, 1, 2):
    ctx.set_option(L.OPT_RAY_SORT, sort)
    ctx.cast_device(d_rays.value, n, d_hits.value, depth)
    ctx.sync()
    ctx.timer_begin()
    for _ in range(5):
        ctx.cast_device(d_rays.value, n, d_hits.value, depth)
    ms = ctx.timer_end() / 5
    assert rt.cudaMemcpy(hits.ctypes.data_as(C.c_void_p), d_hits, n * 16, 2) == 0
    if ref is None:
        ref = hits.copy()
    print("size %d, %d rays, SVO_OPT_RAY_SORT %d: %.3f ms = %.0f Mrays/s (device-resident, binning included), hit %.2f, mean iter %.1f, identical %s"
          % (size, n, sort, ms, n / ms / 1e3, float((hits["id"] != 0xFFFFFFFF).mean()), float(hits["iter"].mean()), bool(np.array_equal(hits, ref))), flush=True)
