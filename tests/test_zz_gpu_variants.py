"""GPU parity of kernel variants 7 / 8 (tile kernel + warp-local lane refill for the casts after the first) and 9..12 (parent
stack in shared memory, 16-byte stack entries, 72 / 80 registers per thread), through the C ABI against the oracle.  The same kernels run in the CPU suite on the SIMT emulator (test_hostemu.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PLANES = ("rgba8", "depth", "radiance", "hit_id", "iter", "primary_t")


def _planes(c):
    return {"rgba8": c.read_color_rgba8(), "depth": c.read_depth(), "radiance": c.read_radiance(), "hit_id": c.read_hit_id(),
            "iter": c.read_iter(), "primary_t": c.read_primary_t()}


def _assert_equal(got, want, what, planes=PLANES):
    for k in planes:
        g, w = got[k], want[k]
        same = ((g.view(np.uint32) == w.view(np.uint32)) | (np.isnan(g) & np.isnan(w))) if g.dtype.kind == "f" else (g == w)
        assert same.all(), "%s: plane %s differs in %d of %d elements" % (what, k, int((~same).sum()), same.size)


import os

VARIANTS = {7: "refill4", 8: "refill2", 9: "smemstack", 10: "widestack", 11: "regs72", 12: "regs80",
            13: "balanced",  # loop integer work on the FMA pipe (inline PTX)
            14: "wide_bands",
            15: "split",  # aux planes on: falls back to the default kernel; its own path is the production instance below
            16: "split_presetup",
            17: "tile_queue",  # variant 13 as persistent warps that take whole tiles from a queue
            19: "balanced_mask7",  # aux planes on: runs variant 13; the production instance is the ablation
            25: "balanced_9ctas", 26: "balanced_10ctas"}  # 56 / 48 registers (aux planes on: variant 13)


@pytest.mark.parametrize("kernel", list(VARIANTS), ids=list(VARIANTS.values()))
def test_kernel_variants_bit_exact(svo, oracle, terrain512, terrain128, kernel):
    with svo.SvoContext(640, 360) as c:
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.upload(terrain512)
        for cam in ("A", "B", "C"):
            for mode in (0, 2, 1, 3, 4):
                pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
                want, _ = oracle.render(terrain512, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=3, render_mode=mode), 640, 360, nthreads=8)
                c.set_option(svo._lib.OPT_AUX_PLANES, 1)
                c.render(svo.camera_frame(cam, frame_number=3, render_mode=mode))
                _assert_equal(_planes(c), want, "kernel %d cam %s mode %d" % (kernel, cam, mode))
                c.set_option(svo._lib.OPT_AUX_PLANES, 0)  # production instance: content box on
                c.render(svo.camera_frame(cam, frame_number=3, render_mode=mode))
                _assert_equal({"rgba8": c.read_color_rgba8(), "depth": c.read_depth()}, want, "kernel %d cam %s mode %d production" % (kernel, cam, mode),
                              planes=("rgba8", "depth"))
    W, H = 200, 120  # not a multiple of the warp's 16x8 / 16x4 pixels; multi-chunk tree; deeper paths; row bands
    with svo.SvoContext(W, H) as c:
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        c.upload(terrain128)
        for cam, casts in (("A", 2), ("B", 4), ("C", 3)):
            pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
            kw = dict(frame_number=1, render_mode=0, max_depth=7, casts=casts)
            want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, **kw), W, H, nthreads=8)
            f = svo.camera_frame(cam, **kw)
            for y0, y1 in ((0, 37), (37, 38), (38, H)):
                c.render(f, y0, y1)
            _assert_equal(_planes(c), want, "kernel %d terrain128 cam %s casts %d" % (kernel, cam, casts))


def test_persistent_stream_kernel_bit_exact(svo, oracle, terrain512):
    """SVO_OPT_STREAM_KERNEL 1 (persistent threads, warp-level ray fetch) against the oracle and the grid-stride kernel."""
    rng = np.random.default_rng(42)
    n = 300000
    rays = np.zeros(n, dtype=svo.RAY_DTYPE)
    rays["o"] = rng.uniform(0.9, 2.1, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["d"][:50, 0] = 0.0
    rays["d"][50:60] = 0.0
    rays["d"][60:70] = np.nan
    want, _ = oracle.cast_rays(terrain512, rays, max_depth=9, nthreads=8)
    with svo.SvoContext(64, 64) as c:
        c.upload(terrain512)
        for stream_kernel, sort in ((1, 0), (1, 1), (2, 1)):
            c.set_option(svo._lib.OPT_STREAM_KERNEL, stream_kernel)
            c.set_option(svo._lib.OPT_RAY_SORT, sort)
            got = c.cast(rays, 9)
            for k in ("id", "value", "iter"):
                assert np.array_equal(got[k], want[k]), (k, stream_kernel, sort)
            assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), (stream_kernel, sort)


@pytest.mark.parametrize("kernel", [15, 16], ids=["split", "split_presetup"])
def test_split_kernels_bit_exact(svo, oracle, terrain512, terrain128, kernel):
    """Kernel variants 15 / 16 (k_split_primary + k_split_bounce) on the device: colour and depth against the oracle."""
    with svo.SvoContext(640, 360) as c:
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.upload(terrain512)
        for cam in ("A", "B", "C"):
            for casts, mirror in ((2, 0), (3, 0), (4, 1)):
                pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
                kw = dict(frame_number=3, render_mode=0, casts=casts, mirror_value=mirror)
                want, _ = oracle.render(terrain512, oracle.make_frame(pos, l1, l2, r1, r2, **kw), 640, 360, nthreads=8, planes=("rgba8", "depth"))
                n0 = c.launch_count()
                c.render(svo.camera_frame(cam, **kw))
                assert c.launch_count() - n0 == 2
                _assert_equal({"rgba8": c.read_color_rgba8(), "depth": c.read_depth()}, want, "split cam %s casts %d" % (cam, casts), planes=("rgba8", "depth"))
    W, H = 200, 120
    with svo.SvoContext(W, H) as c:
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.upload(terrain128)
        pos, l1, l2, r1, r2 = svo.CAMERAS["C"]
        kw = dict(frame_number=1, render_mode=0, max_depth=7)
        want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, **kw), W, H, nthreads=8, planes=("rgba8", "depth"))
        f = svo.camera_frame("C", **kw)
        for y0, y1 in ((0, 37), (37, 38), (38, H)):
            c.render(f, y0, y1)
        _assert_equal({"rgba8": c.read_color_rgba8(), "depth": c.read_depth()}, want, "split bands", planes=("rgba8", "depth"))
