#!/usr/bin/env python
"""Conservative beam pre-pass (svo_beam_conservative + SVO_FRAME_BEAM_FLOOR) against the plain frame on the bench workload:
kernel time with CUDA events, executed loop iterations, and that the frame does not change.
usage: python tools/beam_bench.py [size=8192] [mode=0]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import svo_raytracer_b200 as svo
from svo_raytracer_b200 import _lib as L

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W, H = 1920, 1080
depth = min(13, int(np.log2(size)))
hm, mm = svo.terrain_inputs(size)
ctx = svo.SvoContext(W, H)
ctx.build_terrain_device(hm, mm, size, min(size, 1024))
out = {"size": size, "mode": mode, "cameras": {}}
for cam in "ABC":
    f = svo.camera_frame(cam, frame_number=1, render_mode=mode, max_depth=depth)
    fb = svo.camera_frame(cam, frame_number=1, render_mode=mode, max_depth=depth, flags=2)
    ctx.render(f)
    rgba, dep = ctx.read_color_rgba8(), ctx.read_depth()
    s0 = ctx.render_stats_executed(f)
    ctx.beam_conservative(fb)
    ctx.render(fb)
    same = bool(np.array_equal(ctx.read_color_rgba8(), rgba) and np.array_equal(ctx.read_depth().view(np.uint32), dep.view(np.uint32)))
    s1 = ctx.render_stats_executed(fb)
    beam = ctx.read_plane(L.PLANE_BEAM)

    def timed(fn, reps=10):
        fn(); ctx.sync(); ctx.timer_begin()
        for _ in range(reps):
            fn()
        return ctx.timer_end() / reps
    t_plain = timed(lambda: ctx.render(f))
    t_pre = timed(lambda: ctx.beam_conservative(fb))
    t_fine = timed(lambda: ctx.render(fb))
    out["cameras"][cam] = {"frame_unchanged": same, "plain_ms": t_plain, "prepass_ms": t_pre, "fine_ms": t_fine, "beam_total_ms": t_pre + t_fine,
                           "iters_plain": s0["iters"], "iters_with_floor": s1["iters"], "iters_saved_pct": 100.0 * (1 - s1["iters"] / max(1, s0["iters"])),
                           "blocks_all_miss_pct": 100.0 * float(np.isinf(beam).mean()), "casts": s0["casts"]}
    print(cam, json.dumps(out["cameras"][cam]), flush=True)
print(json.dumps(out))
