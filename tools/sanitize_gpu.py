#!/usr/bin/env python
"""One pass over every kernel family of libsvo_b200.so on small inputs, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 9 python tools/sanitize_gpu.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_gpu.py
    compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_gpu.py

(tools/gpu_sanitize.sh runs the three and keeps the summaries for profiles/).  No torch, no oracle: ctypes + numpy only, so
that the sanitizer sees nothing but this library's launches.  Every result is still compared with the default kernel's."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import svo_raytracer_b200 as svo  # noqa: E402
from svo_raytracer_b200 import _lib as L  # noqa: E402

QUICK = "--quick" in sys.argv
W, H, N = (160, 96, 64) if QUICK else (320, 184, 128)
DEPTH = int(np.log2(N))
t0 = time.time()


def say(what):
    print("[%6.1f s] %s" % (time.time() - t0, what), flush=True)


hm, mm = svo.terrain_inputs(N)
host = svo.build_terrain(hm, mm, N, N // 2)
ctx = svo.SvoContext(W, H)

# world generation on the device + whole transcode (k_pyramid, k_classify, k_expand, k_sizes, k_offsets, k_emit, k_transcode_level)
nbytes = ctx.build_terrain_device(hm, mm, N, N // 2)
dev = ctx.download()
assert nbytes == host.size and np.array_equal(dev, host), "device builder differs from the host builder"
say("device builder: %d bytes, byte-equal" % nbytes)
ctx.upload(host)

frames = {m: svo.camera_frame("B", frame_number=2, render_mode=m, max_depth=DEPTH) for m in range(5)}
frames["C0"] = svo.camera_frame("C", frame_number=3, render_mode=0, max_depth=DEPTH, casts=3)
want = {}
for kernel in (13, 17, 0, 10, 9, 4, 6, 7, 1, 2, 14, 15, 16, 19):
    ctx.set_option(L.OPT_KERNEL, kernel)
    for aux in (0, 1):
        ctx.set_option(L.OPT_AUX_PLANES, aux)
        for key, f in frames.items():
            ctx.render(f)
            got = (ctx.read_color_rgba8(), ctx.read_depth())
            if kernel == 13 and aux == 0:
                want[key] = got
            assert np.array_equal(got[0], want[key][0]) and np.array_equal(got[1].view(np.uint32), want[key][1].view(np.uint32)), (kernel, aux, key)
    say("render kernel %d ok" % kernel)
ctx.set_option(L.OPT_AUX_PLANES, 0)
ctx.set_option(L.OPT_KERNEL, 13)

# row bands, interleaved bands with the in-kernel fence, lanes, pipelined read-back
f0 = frames[0]
for y0, y1 in ((0, 37), (37, 38), (38, H)):
    ctx.render(f0, y0, y1)
assert np.array_equal(ctx.read_color_rgba8(), want[0][0])
fence = ctx.fence_device_ptr()
for parts in (1, 3):
    ctx.fence_reset()
    for part in range(parts):
        ctx.render_interleaved_signal(f0, part, parts, (fence,), slot=2)
    ctx.fence_wait(parts, slot=2)
    ctx.fence_wait_signal(parts, 2, (fence,), 3)
    ctx.fence_wait(1, slot=3)
    ctx.sync()
    assert np.array_equal(ctx.read_color_rgba8(), want[0][0]), parts
color_h = np.zeros((H, W, 4), np.uint8)
depth_h = np.zeros((H, W), np.float32)
for lane in (0, 1, 2, 6):
    ctx.select_lane(lane)
    ctx.beam_conservative(f0)
    fb = svo.camera_frame("B", frame_number=2, render_mode=0, max_depth=DEPTH, flags=2)
    ctx.render(fb)
    ctx.read_planes_async(color_h.ctypes.data, depth_h.ctypes.data)
    ctx.read_wait()
    assert np.array_equal(color_h, want[0][0]) and np.array_equal(depth_h.view(np.uint32), want[0][1].view(np.uint32)), lane
ctx.select_lane(0)
say("bands, fences, lanes, beam floor ok")

# beam pre-passes: as shipped, conservative, conservative in row parts
ctx.beam(f0)
ctx.beam_conservative(f0)
beam = ctx.read_plane(L.PLANE_BEAM).copy()
lh = H // 4 + 1
ctx.fence_reset()
own = ctx.device_ptr(L.PLANE_BEAM_LATTICE)
for r in range(3):
    ctx.beam_lattice_rows(f0, r * lh // 3, (r + 1) * lh // 3, (own,), (fence,), slot=9)
ctx.fence_wait(3, slot=9)
ctx.beam_filter()
assert np.array_equal(ctx.read_plane(L.PLANE_BEAM).view(np.uint32), beam.view(np.uint32))
say("beam pre-passes ok")

# ray streams: grid-stride / persistent / wide-stack kernels, with and without binning (CUB sort)
rng = np.random.default_rng(1)
n = 20000 if QUICK else 70000
rays = np.zeros(n, dtype=svo.RAY_DTYPE)
rays["o"] = rng.uniform(0.9, 2.1, (n, 3)).astype(np.float32)
d = rng.normal(size=(n, 3))
rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
rays["d"][:8] = 0.0
rays["d"][8:16] = np.nan
first = None
for sk, sort in ((0, 0), (0, 1), (0, 2), (1, 2), (2, 1)):
    ctx.set_option(L.OPT_STREAM_KERNEL, sk)
    ctx.set_option(L.OPT_RAY_SORT, sort)
    got = ctx.cast(rays, DEPTH)
    if first is None:
        first = got
    for k in ("id", "value", "iter"):
        assert np.array_equal(got[k], first[k]), (sk, sort, k)
    assert np.array_equal(got["t"].view(np.uint32), first["t"].view(np.uint32)), (sk, sort)
ctx.set_option(L.OPT_STREAM_KERNEL, 0)
say("ray streams ok")

# instrumented kernels, probes
ctx.render_stats(f0)
ctx.render_stats_executed(f0)
ctx.gather_probe(1 << 22, 32)
x = rng.uniform(-4, 4, 4096).astype(np.float32)
for fn in range(4):
    ctx.math_probe(fn, np.clip(x, -1, 1) if fn == 2 else x)
say("stats and probes ok")

# incremental edit (k_diff_apply, k_mark_dirty, k_select_roots, the level kernels on sub-trees), then the whole transcode again
edited = host.copy()
cp = int.from_bytes(edited[1:5].tobytes(), "big")
hits = [i for i in range(8) if edited[cp + 7 * i] != 0] if ((int(edited[5]) << 8) | int(edited[6])) == 0 else []
off = cp + 7 * hits[0] if hits else cp
edited[off] = edited[off] % 3 + 1 if edited[off] else 0
ctx.upload_range(edited, off, off + 7)
canon_inc = ctx.scene_canonical()
canon_inc.pop("stored")  # a patched array also holds the replaced sub-trees until the next whole transcode
ctx.render(f0)
inc = ctx.read_color_rgba8()
ctx.upload(edited)
canon_whole = ctx.scene_canonical()
canon_whole.pop("stored")
assert canon_whole == canon_inc, (canon_whole, canon_inc)
ctx.render(f0)
assert np.array_equal(ctx.read_color_rgba8(), inc)
say("incremental transcode ok")
ctx.close()
print("sanitize sequence ok")
