#!/usr/bin/env python
"""Developer micro-bench: one scene, several kernel variants / options, per-camera kernel time (CUDA events).
usage: python tools/kbench.py [size] [variants e.g. 0,2,0f,2f]   (suffix f = fast math)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import svo_raytracer_b200 as svo
from svo_raytracer_b200 import _lib as L

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
variants = (sys.argv[2] if len(sys.argv) > 2 else "0,2").split(",")
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
W, H = 1920, 1080
t0 = time.time()
cache = "/tmp/kbench_world_%d.npy" % size  # several libraries (SVO_B200_LIB=...) are compared on one box: build once
if os.path.exists(cache):
    nodes = np.load(cache)
else:
    hm, mm = svo.terrain_inputs(size)
    nodes = svo.build_terrain(hm, mm, size, min(size, 1024))
    np.save(cache, nodes)
print("world %d^3: %.1f MB in %.1fs" % (size, nodes.size / 1e6, time.time() - t0), flush=True)
depth = min(13, int(np.log2(size)))
ctx = svo.SvoContext(W, H)
ctx.upload(nodes)
stats = {c: ctx.render_stats(svo.camera_frame(c, frame_number=1, render_mode=mode, max_depth=depth)) for c in "ABC"}
for v in variants:
    fast = "f" in v
    k = int(v.rstrip("fnp"))
    if "p" in v:  # L2 access-policy window over the descriptor prefix (takes effect at upload)
        ctx.set_option(L.OPT_L2_PERSIST, 1)
        ctx.upload(nodes)
    ctx.set_option(L.OPT_KERNEL, k)
    ctx.set_option(L.OPT_FAST_MATH, int(fast))
    ctx.set_option(L.OPT_CONTENT_BOUNDS, 0 if "n" in v else 1)
    line = []
    tot_ms, tot_rays = 0.0, 0
    for cam in "ABC":
        frames = [svo.camera_frame(cam, frame_number=i + 1, render_mode=mode, max_depth=depth) for i in range(12)]
        for f in frames[:2]:
            ctx.render(f)
        ctx.sync()
        ctx.timer_begin()
        for f in frames[2:]:
            ctx.render(f)
        ms = ctx.timer_end() / 10
        rays = stats[cam]["casts"]
        line.append("%s %.3fms %.0fMr/s" % (cam, ms, rays / ms / 1e3))
        tot_ms += ms
        tot_rays += rays
    print("variant %-3s | %s | cycle %.3f ms/frame %.0f Mrays/s" % (v, " | ".join(line), tot_ms / 3, tot_rays / tot_ms / 1e3), flush=True)
