#!/usr/bin/env python
"""Incoherent secondary-ray stress (BASELINE configs[3] in spirit): N random-direction rays starting just above
the terrain surface, traced through the ray-stream API on device-resident buffers, with and without ray binning.
usage: python tools/stream_bench.py [size=8192] [n_rays=16777216]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import svo_raytracer_b200 as svo
from svo_raytracer_b200 import _lib as L

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
hm, mm = svo.terrain_inputs(size)
nodes = svo.build_terrain(hm, mm, size, min(size, 1024))
ctx = svo.SvoContext(64, 64)
ctx.upload(nodes)
depth = min(13, int(np.log2(size)))
g = torch.Generator(device="cuda").manual_seed(42)
xz = torch.randint(0, size, (n, 2), device="cuda", generator=g)
h = torch.from_numpy(((hm.astype(np.uint32) * (size // 4)) >> 16).astype(np.int32)).cuda()
y = h[xz[:, 1], xz[:, 0]].float() + 2.5
o = torch.stack([(xz[:, 0].float() + 0.5) / size + 1.0, y / size + 1.0, (xz[:, 1].float() + 0.5) / size + 1.0], 1)
d = torch.randn((n, 3), device="cuda", generator=g)
d = d / d.norm(dim=1, keepdim=True)
rays = torch.cat([o, d], 1).contiguous()
hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ref = None
for sort, stream_kernel in ((0, 0), (1, 0), (0, 1), (1, 1), (1, 2), (1, 0), (1, 1), (1, 2)):  # stream_kernel 1 = persistent threads with warp-level ray fetch, 2 = grid-stride with 16-byte stack entries
    ctx.set_option(L.OPT_RAY_SORT, sort)
    ctx.set_option(L.OPT_STREAM_KERNEL, stream_kernel)
    ctx.cast_device(rays.data_ptr(), n, hits.data_ptr(), depth)
    torch.cuda.synchronize()
    ctx.timer_begin()
    for _ in range(3):
        ctx.cast_device(rays.data_ptr(), n, hits.data_ptr(), depth)
    ms = ctx.timer_end() / 3
    if ref is None:
        ref = hits.clone()
    same = bool(torch.equal(ref, hits))
    hit_frac = float((hits[:, 0] != -1).float().mean())
    print("size %d rays %d sort %d stream kernel %d: %.2f ms  %.0f Mrays/s  hit %.2f  mean iter %.1f  identical %s" % (
        size, n, sort, stream_kernel, ms, n / ms / 1e3, hit_frac, float(hits[:, 3].float().mean()), same), flush=True)
