"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs.  Bit-exact for ids / iteration counts / depth /
rgba8 (integer and index work, and -- because the arithmetic contract is fixed
on both sides -- the float planes too)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PLANES = ("rgba8", "depth", "radiance", "hit_id", "iter", "primary_t")


def _render_gpu(svo, ctx, frame):
    ctx.render(frame)
    return {"rgba8": ctx.read_color_rgba8(), "depth": ctx.read_depth(), "radiance": ctx.read_radiance(),
            "hit_id": ctx.read_hit_id(), "iter": ctx.read_iter(), "primary_t": ctx.read_primary_t()}


def _assert_planes_equal(got, want, what):
    for k in PLANES:
        g, w = got[k], want[k]
        if g.dtype.kind == "f":
            same = (g.view(np.uint32) == w.view(np.uint32)) | (np.isnan(g) & np.isnan(w))
        else:
            same = g == w
        bad = int((~same).sum())
        assert bad == 0, "%s: plane %s differs in %d of %d elements" % (what, k, bad, same.size)


KERNELS = [0, 1, 2, 6]  # SVO_OPT_KERNEL: 0 tile, 1 persistent megakernel, 2 wavefront, 6 tile + per-CTA octant binning of bounce rays


@pytest.fixture(scope="module", params=KERNELS, ids=["tile", "persistent", "wavefront", "binned"])
def ctx512(request, svo, terrain512):
    c = svo.SvoContext(640, 360)
    c.set_option(svo._lib.OPT_AUX_PLANES, 1)
    c.set_option(svo._lib.OPT_KERNEL, request.param)
    c.upload(terrain512)
    yield c
    c.close()


@pytest.mark.parametrize("cam", ["A", "B", "C"])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_config1_all_modes_bit_exact(svo, oracle, terrain512, ctx512, cam, mode):
    """BASELINE configs[0]: 512^3 terrain, 640x360, every render mode, cameras A/B/C."""
    pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
    want, st = oracle.render(terrain512, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=3, render_mode=mode),
                             640, 360, nthreads=8)
    assert st.stale_pops == 0
    got = _render_gpu(svo, ctx512, svo.camera_frame(cam, frame_number=3, render_mode=mode))
    _assert_planes_equal(got, want, "cam %s mode %d" % (cam, mode))


def test_math_probe_bit_exact(svo, oracle, ctx512):
    rng = np.random.default_rng(7)
    L = oracle.lib()
    cases = {
        0: np.concatenate([rng.uniform(-7, 7, 20000), rng.uniform(-2e6, 2e6, 20000), [0.0, -0.0, np.inf, np.nan, 1e9, 3e38]]),
        1: np.concatenate([rng.uniform(-7, 7, 20000), rng.uniform(-2e6, 2e6, 20000), [0.0, -0.0, np.inf, np.nan]]),
        2: np.concatenate([rng.uniform(-1, 1, 40000), [-1.0, 1.0, 0.5, -0.5, 1.0000001, np.nan]]),
        3: np.concatenate([rng.uniform(-20, 5, 40000), rng.uniform(-110, 90, 2000), [0.0, np.nan, -np.inf, np.inf]]),
    }
    names = {0: "sin", 1: "cos", 2: "acos", 3: "exp"}
    for fn, x in cases.items():
        x = x.astype(np.float32)
        got = ctx512.math_probe(fn, x)
        f = getattr(L, "svo_oracle_" + names[fn])
        want = np.array([f(float(v)) for v in x], dtype=np.float32)
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), "%s differs at %s" % (names[fn], x[~same][:5])
    x = rng.uniform(0, 8192, 20000).astype(np.float32)
    y = rng.uniform(0, 8192, 20000).astype(np.float32)
    got = ctx512.math_probe(4, x, y)
    want = np.array([L.svo_oracle_rand(float(a), float(b)) for a, b in zip(x, y)], dtype=np.float32)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()


def test_ray_stream_bit_exact(svo, oracle, terrain512, ctx512):
    """Incoherent rays (BASELINE configs[3] in miniature): random origins in/around the cube, random directions,
    including axis-parallel, zero and NaN directions."""
    rng = np.random.default_rng(42)
    n = 200000
    rays = np.zeros(n, dtype=svo.RAY_DTYPE)
    rays["o"] = rng.uniform(0.9, 2.1, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["d"][:50, 0] = 0.0          # axis-parallel (t_coef = -inf quirk)
    rays["d"][50:60] = 0.0           # zero direction
    rays["d"][60:70] = np.nan        # NaN direction (zero-normal bounce)
    rays["d"][70:80, 1] = np.nan
    rays["o"][80:90] = np.nan
    for depth in (13, 9, 5):
        want, st = oracle.cast_rays(terrain512, rays, max_depth=depth, nthreads=8)
        got = ctx512.cast(rays, max_depth=depth)
        for k in ("id", "value", "iter"):
            assert (got[k] == want[k]).all(), "depth %d field %s: %d mismatches" % (depth, k, (got[k] != want[k]).sum())
        assert (got["t"].view(np.uint32) == want["t"].view(np.uint32)).all()
        assert st.stale_pops == 0


@pytest.mark.parametrize("kernel", KERNELS)
def test_multichunk_tree_and_rows(svo, oracle, terrain128, kernel):
    """Fill levels + chunk splices (128^3 world of 64^3 chunks) and the row-band partition."""
    W, H = 200, 120  # not a multiple of the 16x8 CTA tile
    with svo.SvoContext(W, H) as c:
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.upload(terrain128)
        for cam in ("A", "B", "C"):
            pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
            want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=0, max_depth=7),
                                    W, H, nthreads=8)
            f = svo.camera_frame(cam, frame_number=1, render_mode=0, max_depth=7)
            # render as three uneven bands, as the multi-GPU tile partition does
            for y0, y1 in ((0, 37), (37, 38), (38, H)):
                c.render(f, y0, y1)
            got = {"rgba8": c.read_color_rgba8(), "depth": c.read_depth(), "radiance": c.read_radiance(),
                   "hit_id": c.read_hit_id(), "iter": c.read_iter(), "primary_t": c.read_primary_t()}
            _assert_planes_equal(got, want, "terrain128 cam " + cam)


def test_render_stats_match_oracle(svo, oracle, terrain512):
    """The instrumented kernel's counters (casts, iterations, reference-layout record bytes: the algorithmic
    bytes bench.py reports) equal the oracle's."""
    with svo.SvoContext(640, 360) as c:
        c.upload(terrain512)
        for cam, mode in (("A", 0), ("B", 2), ("C", 0)):
            pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
            _, st = oracle.render(terrain512, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=2, render_mode=mode),
                                  640, 360, nthreads=8, planes=("depth",))
            got = c.render_stats(svo.camera_frame(cam, frame_number=2, render_mode=mode))
            assert got == {"casts": st.casts, "iters": st.iters, "record_bytes": st.record_bytes}, (cam, mode, got, st.as_dict())


def test_fast_math_is_close(svo, oracle, terrain512):
    """SVO_OPT_FAST_MATH lets the t arithmetic contract into FFMA: not bit-exact by design; hit ids may flip on
    a handful of silhouette pixels and t moves by a few ulp (north_star tolerance: relative t error <= 1e-5)."""
    with svo.SvoContext(640, 360) as c:
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        c.set_option(svo._lib.OPT_FAST_MATH, 1)
        c.upload(terrain512)
        pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
        want, _ = oracle.render(terrain512, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=3), 640, 360, nthreads=8)
        c.render(svo.camera_frame("B", frame_number=1, render_mode=3))
        ids, t = c.read_hit_id(), c.read_primary_t()
        same = ids == want["hit_id"]
        assert same.mean() > 0.999
        both = same & (ids != svo.NO_HIT)
        rel = np.abs(t[both] - want["primary_t"][both]) / np.maximum(want["primary_t"][both], 1e-6)
        assert np.quantile(rel, 0.999) <= 1e-5 and rel.max() <= 1e-3


def test_renderer_mirror_reads_like_the_engine(svo, oracle, terrain128):
    """The Renderer.java mirror, driven the way Main.preRun / Main.updateEarly drive the GL renderer
    (Main.java:102-122, :267-285, :132-146), including the partial-upload path (Main.java:349-350)."""
    svo.Renderer.resetInstance()
    W, H = 160, 90
    r = svo.Renderer.getInstance(W, H)
    trace = r.addShader("svotrace", "src/shaders/svotrace.comp")
    beam = r.addShader("beamShader", "src/shaders/svobeam.comp")
    assert r.getShaderByName("svotrace") is trace
    buf = np.zeros(terrain128.size + 4096, np.uint8)  # Octree's big ByteBuffer; memOffset bytes are valid
    buf[:terrain128.size] = terrain128
    r.addSSBO(7, buf, terrain128.size)
    pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
    r.useProgram(trace)
    r.glUniform3fv(8, pos)
    for loc, v in zip((1, 2, 3, 4), (l1, l2, r1, r2)):
        r.glUniform3fv(loc, v)
    r.glUniform1i(5, 1)
    r.glUniform1i(6, 2)
    r.glUniform1i(9, terrain128.size)
    r.glUniform1i(11, 0)
    r._frame.maxDepth = 7
    r.dispatchCompute(trace, (W + 7) // 8, (H + 7) // 8, 1)
    assert r.printGLErrors() == 0
    want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=2, max_depth=7), W, H, nthreads=4)
    assert np.array_equal(r.getFramebufferImage(), want["rgba8"])
    depth = r.getDepthImage()
    assert np.array_equal(depth.view(np.uint32), want["depth"].view(np.uint32))
    assert r.ctx.read_depth_at(W // 2, H // 2) == depth[H // 2, W // 2]  # the crosshair pick

    # an "SDF edit": change a leaf value in place and append nothing -> updateSSBO(7, buf, start, end)
    edited = terrain128.copy()
    hits = want["hit_id"][:, 20:]  # outside the 10x10 debug overlay (svotrace.comp:700)
    solid = np.nonzero(hits != svo.NO_HIT)
    ptr = int(hits[solid[0][len(solid[0]) // 2], solid[1][len(solid[0]) // 2]])
    edited[ptr] = 3 if edited[ptr] != 3 else 2
    buf[:edited.size] = edited
    r.updateSSBO(7, buf, ptr, ptr + 1)
    r.updateSSBO(7, buf, 5, 5)  # Renderer.java:137-140: prints "Update SSBO error: Invalid parameters." and returns
    r.dispatchCompute(trace, (W + 7) // 8, (H + 7) // 8, 1)
    want2, _ = oracle.render(edited, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=2, max_depth=7), W, H, nthreads=4)
    assert np.array_equal(r.getFramebufferImage(), want2["rgba8"])
    assert not np.array_equal(want2["rgba8"], want["rgba8"])

    # beam pre-pass (Main.java:257-266) then a trace that consumes it (uniform 11)
    r.useProgram(beam)
    r.dispatchCompute(beam, 1, 1, 1)
    got_beam = r.ctx.read_plane(svo._lib.PLANE_BEAM)
    f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=2, max_depth=7)
    want_beam = oracle.beam(edited, f, W, H)
    assert np.array_equal(got_beam.view(np.uint32), want_beam.view(np.uint32))
    r.glUniform1i(11, 1)
    r.dispatchCompute(trace, 1, 1, 1)
    f.useBeam = 1
    want3, _ = oracle.render(edited, f, W, H, beam=want_beam, nthreads=4)
    assert np.array_equal(r.getFramebufferImage(), want3["rgba8"])
    assert r.printGLErrors() == 0
    svo.Renderer.resetInstance()


def test_errors_are_codes_not_crashes(svo, terrain128):
    with svo.SvoContext(64, 64) as c:
        with pytest.raises(svo.SvoError) as e:
            c.render(svo.camera_frame("A"))
        assert e.value.code == svo._lib.ERR_NO_SCENE
        c.upload(terrain128)
        bad = svo.camera_frame("A", max_depth=40)
        with pytest.raises(svo.SvoError) as e:
            c.render(bad)
        assert e.value.code == svo._lib.ERR_INVALID
        with pytest.raises(svo.SvoError):
            c.render(svo.camera_frame("A"), 10, 200)
        # a stream that is not a tree is refused, and leaves the context without a scene rather than with a stale one
        from test_transcode import cyclic_stream
        with pytest.raises(svo.SvoError) as e:
            c.upload(cyclic_stream())
        assert e.value.code == svo._lib.ERR_FORMAT
        with pytest.raises(svo.SvoError) as e:
            c.render(svo.camera_frame("A"))
        assert e.value.code == svo._lib.ERR_NO_SCENE
        with pytest.raises(svo.SvoError) as e:
            c.set_option(svo._lib.OPT_KERNEL, 3)
        assert e.value.code == svo._lib.ERR_INVALID
        c.upload(terrain128)
        c.render(svo.camera_frame("A"))
        c.upload(np.zeros(0, np.uint8))  # an empty stream is legal: everything misses
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        c.render(svo.camera_frame("A", render_mode=3))
        assert (c.read_hit_id() == svo.NO_HIT).all()


@pytest.mark.parametrize("kernel", [17, 14, 0], ids=["tile_queue", "wide_bands", "tile"])
def test_interleaved_band_partition_single_gpu(svo, oracle, terrain128, kernel):
    """svo_render_interleaved: parts 0..2 of 3 rendered one after the other fill the same frame the oracle renders.
    svo_render_interleaved_signal: the same plus the frame-complete fence, bumped once per launch."""
    W, H = 200, 117
    with svo.SvoContext(W, H) as c:
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        c.upload(terrain128)
        pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
        want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=2, render_mode=0, max_depth=7), W, H, nthreads=4)
        f = svo.camera_frame("B", frame_number=2, render_mode=0, max_depth=7)
        for band_rows, parts in ((8, 3), (24, 2), (64, 3), (32, 8)):
            c.set_option(svo._lib.OPT_BAND_ROWS, band_rows)
            c.render(svo.camera_frame("A", frame_number=1, render_mode=3, max_depth=7))  # scribble over every plane first
            for part in reversed(range(parts)):
                c.render_interleaved(f, part, parts)
            got = {"rgba8": c.read_color_rgba8(), "depth": c.read_depth(), "radiance": c.read_radiance(),
                   "hit_id": c.read_hit_id(), "iter": c.read_iter(), "primary_t": c.read_primary_t()}
            _assert_planes_equal(got, want, "interleaved %d rows x %d parts" % (band_rows, parts))
        # fused signal on the context's own fence (slot 1): after `parts` launches the counter stands at `parts`
        c.set_option(svo._lib.OPT_AUX_PLANES, 0)
        c.set_option(svo._lib.OPT_BAND_ROWS, 8)
        c.render(svo.camera_frame("A", frame_number=1, render_mode=3, max_depth=7))
        for rep in range(2):
            for part in range(5):
                c.render_interleaved_signal(f, part, 5, (), slot=1)
            c.fence_wait(5 * (rep + 1), slot=1)
        c.sync()  # raises SVO_ERR_FENCE if a wait ran into the watchdog
        assert np.array_equal(c.read_color_rgba8(), want["rgba8"]) and np.array_equal(c.read_depth().view(np.uint32), want["depth"].view(np.uint32))


def test_two_lanes_overlap_frames_without_mixing_them(svo, oracle, terrain512):
    """svo_select_lane: frames rendered back to back on alternating lanes (two streams, two plane sets) may overlap on the
    GPU; each lane must end up holding exactly the last frame drawn on it, and the pipelined read-back path
    (svo_read_planes_async + svo_swap_buffers, which alternates lanes) must deliver every frame intact."""
    W, H = 640, 360
    cams = ["A", "B", "C", "C", "B", "A", "C"]
    want = []
    for i, cam in enumerate(cams):
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        p, _ = oracle.render(terrain512, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=i + 1, render_mode=0), W, H, nthreads=8,
                             planes=("rgba8", "depth"))
        want.append(p)
    for kernel in (13, 17, 0):
        with svo.SvoContext(W, H) as c:
            c.set_option(svo._lib.OPT_KERNEL, kernel)
            c.upload(terrain512)
            c.timer_begin()
            for i, cam in enumerate(cams):
                c.select_lane(i & 1)
                c.render(svo.camera_frame(cam, frame_number=i + 1, render_mode=0))
            assert c.timer_end() > 0
            for lane, i in ((0, 6), (1, 5)):
                c.select_lane(lane)
                assert np.array_equal(c.read_color_rgba8(), want[i]["rgba8"]), (kernel, lane)
                assert np.array_equal(c.read_depth().view(np.uint32), want[i]["depth"].view(np.uint32)), (kernel, lane)
            # three lanes, each frame preceded by its own conservative beam pre-pass (per-lane beam planes): nothing mixes
            for i, cam in enumerate(cams[:6]):
                c.select_lane(i % 3)
                fb = svo.camera_frame(cam, frame_number=i + 1, render_mode=0, flags=2)
                c.beam_conservative(fb)
                c.render(fb)
            for lane, i in ((0, 3), (1, 4), (2, 5)):
                c.select_lane(lane)
                assert np.array_equal(c.read_color_rgba8(), want[i]["rgba8"]), (kernel, "beam", lane)
                assert np.array_equal(c.read_depth().view(np.uint32), want[i]["depth"].view(np.uint32)), (kernel, "beam", lane)
            c.select_lane(0)
            # pipelined read-back: every frame lands in its host buffer
            import torch
            bufs = [(torch.empty((H, W, 4), dtype=torch.uint8).pin_memory(), torch.empty((H, W), dtype=torch.float32).pin_memory()) for _ in cams]
            for i, cam in enumerate(cams):
                c.render(svo.camera_frame(cam, frame_number=i + 1, render_mode=0))
                c.read_planes_async(bufs[i][0].data_ptr(), bufs[i][1].data_ptr())
                c.swap_buffers()
            c.read_wait()
            for i in range(len(cams)):
                assert np.array_equal(bufs[i][0].numpy(), want[i]["rgba8"]), (kernel, i)
                assert np.array_equal(bufs[i][1].numpy().view(np.uint32), want[i]["depth"].view(np.uint32)), (kernel, i)


def test_conservative_beam_full_size(svo, oracle):
    """svo_beam_conservative at 1920x1080 on the 2048^3 world: every block's bound is below the primary hit distance of its
    16 pixels, blocks marked +inf contain no hit, the frame rendered with SVO_FRAME_BEAM_FLOOR equals the frame without
    it bit for bit (colour and depth), and primary iterations are saved."""
    size, W, H = 2048, 1920, 1080
    hm, mm = svo.terrain_inputs(size)
    with svo.SvoContext(W, H) as c:
        c.build_terrain_device(hm, mm, size, 1024)
        for cam in ("A", "B", "C"):
            c.set_option(svo._lib.OPT_AUX_PLANES, 1)
            f3 = svo.camera_frame(cam, frame_number=1, render_mode=3, max_depth=11)
            c.render(f3)
            ids, t = c.read_hit_id(), c.read_primary_t()
            c.set_option(svo._lib.OPT_AUX_PLANES, 0)
            c.beam_conservative(f3)
            beam = c.read_plane(svo._lib.PLANE_BEAM)
            tt = np.where(ids != svo.NO_HIT, t, np.inf).astype(np.float32)
            tmin = tt.reshape(H // 4, 4, W // 4, 4).min(axis=(1, 3))
            assert (beam <= tmin).all(), "cam %s: %d blocks not conservative" % (cam, int((beam > tmin).sum()))
            hit = np.isfinite(tmin)
            if hit.any():
                assert np.median(beam[hit] / tmin[hit]) > 0.9, float(np.median(beam[hit] / tmin[hit]))
            for mode in (0, 2):
                f = svo.camera_frame(cam, frame_number=2, render_mode=mode, max_depth=11)
                c.render(f)
                rgba, depth = c.read_color_rgba8(), c.read_depth()
                st0 = c.render_stats_executed(f)
                fb = svo.camera_frame(cam, frame_number=2, render_mode=mode, max_depth=11, flags=2)
                c.render(fb)
                assert np.array_equal(c.read_color_rgba8(), rgba), (cam, mode)
                assert np.array_equal(c.read_depth().view(np.uint32), depth.view(np.uint32)), (cam, mode)
                st1 = c.render_stats_executed(fb)
                if mode == 0:
                    assert st1["iters"] < st0["iters"], (cam, mode, st0, st1)
                else:  # mode 2 reads the primary's stale iteration count (penumbra, svotrace.comp:616-619): the floor is ignored there
                    assert st1["iters"] == st0["iters"], (cam, mode, st0, st1)


def test_beam_lattice_in_row_parts_equals_the_whole_prepass(svo, oracle, terrain128):
    """svo_beam_lattice_rows + svo_beam_filter (the tile partition's shared pre-pass: every rank traces a slice of the lattice)
    produce the beam plane of svo_beam_conservative bit for bit, whatever the slicing; the launch's last CTA bumps the fences."""
    W, H = 256, 144
    L = svo._lib
    with svo.SvoContext(W, H) as c:
        c.upload(terrain128)
        f = svo.camera_frame("A", frame_number=1, render_mode=0, max_depth=7)
        c.beam_conservative(f)
        want = c.read_plane(L.PLANE_BEAM).copy()
        lh = H // 4 + 1
        own = c.device_ptr(L.PLANE_BEAM_LATTICE)
        fence = c.fence_device_ptr()
        assert own and fence
        for parts in (1, 2, 3, 8):
            c.beam_lattice_rows(f, 0, lh, (), (), 0)  # overwritten below part by part
            c.fence_reset()
            for r in reversed(range(parts)):
                c.beam_lattice_rows(f, r * lh // parts, (r + 1) * lh // parts, (own,), (fence,), slot=9)
            c.fence_wait(parts, slot=9)
            c.beam_filter()
            c.sync()
            got = c.read_plane(L.PLANE_BEAM)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), parts
        c.fence_reset()
        c.beam_lattice_rows(f, 5, 3, (), (fence,), slot=9)  # an empty slice still owes its bump
        c.fence_wait(1, slot=9)
        c.sync()


@pytest.mark.parametrize("band_rows,parts", [(8, 3), (16, 2), (8, 8), (64, 3)])
def test_read_interleaved_assembles_the_frame_in_host_memory(svo, oracle, terrain128, band_rows, parts):
    """svo_read_interleaved_async: each part's bands go from the context's planes to their place in full-frame HOST buffers
    (one strided copy per plane; the last band is cut by the image's bottom edge: H = 117).  Parts rendered and read one
    after the other, on rotating lanes, assemble the oracle's frame; rows of other parts are not touched."""
    import torch
    W, H = 200, 117
    pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
    want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=2, render_mode=0, max_depth=7), W, H, nthreads=4,
                            planes=("rgba8", "depth"))
    f = svo.camera_frame("B", frame_number=2, render_mode=0, max_depth=7)
    color = torch.full((H, W, 4), 7, dtype=torch.uint8).pin_memory()
    depth = torch.full((H, W), -5.0, dtype=torch.float32).pin_memory()
    with svo.SvoContext(W, H) as c:
        c.upload(terrain128)
        c.set_option(svo._lib.OPT_BAND_ROWS, band_rows)
        for part in range(parts):
            c.select_lane(part % 7)
            c.render_interleaved(f, part, parts)
            c.read_interleaved_async(part, parts, color.data_ptr(), depth.data_ptr())
            if part == 0:  # only part 0's bands have arrived
                c.read_wait()
                rows = np.zeros(H, bool)
                for b in range(0, (H + band_rows - 1) // band_rows, parts):
                    rows[b * band_rows:(b + 1) * band_rows] = True
                assert np.array_equal(color.numpy()[rows], want["rgba8"][rows])
                if parts > 1:
                    assert (color.numpy()[~rows] == 7).all() and (depth.numpy()[~rows] == -5.0).all()
        c.read_wait()
        assert np.array_equal(color.numpy(), want["rgba8"])
        assert np.array_equal(depth.numpy().view(np.uint32), want["depth"].view(np.uint32))
        with pytest.raises(svo.SvoError):
            c.select_lane(7)


def test_fence_watchdog_is_reported(svo, terrain128):
    """A wait nobody signals gives up after ~2 s; svo_sync reports it (the latch was invisible to the host)."""
    with svo.SvoContext(64, 64) as c:
        c.upload(terrain128)
        c.fence_wait(1, slot=3)
        with pytest.raises(svo.SvoError) as e:
            c.sync()
        assert e.value.code == svo._lib.ERR_FENCE
        c.fence_reset()
        c.fence_signal((), slot=3)
        c.fence_wait(1, slot=3)
        c.sync()


def _ipc_worker(rank, world, port, q):
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import svo_raytracer_b200 as svo
        from svo_raytracer_b200 import _lib as L
        W, H, n = 320, 200, 128
        hm, mm = svo.terrain_inputs(n)
        nodes = svo.build_terrain(hm, mm, n, 64)  # replicated octree
        ctx = svo.SvoContext(W, H, device=rank)
        ctx.upload(nodes)
        planes = (L.PLANE_COLOR_RGBA8, L.PLANE_DEPTH)
        handles = [ctx.ipc_export(p) for p in planes] if rank == 0 else [None, None]
        dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            for p, h in zip(planes, handles):
                ctx.bind_plane(p, ctx.ipc_import(h))
        # frame-complete fence without a collective: counters bumped by remote atomics over NVLink
        fh = [None] * world
        dist.all_gather_object(fh, ctx.fence_export())
        f = svo.camera_frame("B", frame_number=1, render_mode=2, max_depth=7)
        peers = [ctx.ipc_import(fh[r]) for r in range(1, world)] if rank == 0 else None
        owner = [ctx.ipc_import(fh[0])] if rank != 0 else None
        out = {}
        k = 0
        # 1) tile kernel + separate fence-signal launches; 2) the tile-queue kernel (variant 17) whose last CTA bumps the
        # frame-complete fence itself: one launch per rank and frame.  Three frames each: the peer may not overwrite a
        # frame the owner has not consumed.
        for tag, kernel in (("separate", 0), ("fused", 17)):
            ctx.set_option(L.OPT_KERNEL, kernel)
            for _ in range(3):
                if rank == 0:
                    if tag == "fused":
                        ctx.render_interleaved_signal(f, 0, world)
                    else:
                        ctx.render_interleaved(f, 0, world)
                        ctx.fence_signal()
                    ctx.fence_wait((k + 1) * world)
                    out[tag] = (ctx.read_color_rgba8(), ctx.read_depth())  # consume (stream-ordered after the wait)
                    ctx.fence_signal(peers)
                else:
                    ctx.fence_wait(k)
                    if tag == "fused":
                        ctx.render_interleaved_signal(f, rank, world, owner)
                    else:
                        ctx.render_interleaved(f, rank, world)
                        ctx.fence_signal(owner)
                k += 1
            if rank == 0:  # scribble over the planes between the two passes
                ctx.render(svo.camera_frame("A", frame_number=1, render_mode=3, max_depth=7))
                ctx.sync()
            dist.barrier()
        # 3) the conservative beam pre-pass with the lattice shared between the GPUs: every rank traces its rows of the lattice
        # and stores them into both lattices over NVLink (svo_beam_lattice_rows), waits for everybody's rows, filters
        # locally (svo_beam_filter) and draws its bands with SVO_FRAME_BEAM_FLOOR: same frame, and the same beam plane as
        # svo_beam_conservative computes alone.
        ctx.set_option(L.OPT_KERNEL, 17)
        f0 = svo.camera_frame("B", frame_number=1, render_mode=0, max_depth=7)
        ctx.beam_conservative(f0)
        beam_alone = ctx.read_plane(L.PLANE_BEAM).copy()
        lat_h = [None] * world
        dist.all_gather_object(lat_h, ctx.ipc_export(L.PLANE_BEAM_LATTICE))
        lattices = [ctx.device_ptr(L.PLANE_BEAM_LATTICE) if r == rank else ctx.ipc_import(lat_h[r]) for r in range(world)]
        fences = [ctx.fence_device_ptr()] + (peers if rank == 0 else owner)  # (any order: every listed counter gets the bump)
        lh = H // 4 + 1
        ctx.beam_lattice_rows(f0, 0, lh, (), (), 0)
        ctx.sync()
        dist.barrier()
        fb = svo.camera_frame("B", frame_number=1, render_mode=0, max_depth=7, flags=2)
        same = True
        for i in range(3):
            if rank != 0:
                ctx.fence_wait(k)
            ctx.beam_lattice_rows(fb, rank * lh // world, (rank + 1) * lh // world, lattices, fences, slot=9)
            ctx.fence_wait((i + 1) * world, slot=9)
            ctx.beam_filter()
            if rank == 0:
                ctx.render_interleaved_signal(fb, 0, world)
                ctx.fence_wait((k + 1) * world)
                out["sharedbeam"] = (ctx.read_color_rgba8(), ctx.read_depth())
                same = same and np.array_equal(ctx.read_plane(L.PLANE_BEAM).view(np.uint32), beam_alone.view(np.uint32))
                ctx.fence_signal(peers)
            else:
                ctx.render_interleaved_signal(fb, rank, world, owner)
                same = same and np.array_equal(ctx.read_plane(L.PLANE_BEAM).view(np.uint32), beam_alone.view(np.uint32))
            k += 1
        ctx.sync()
        ok = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        out["beam_planes_equal"] = bool(int(ok.item()))
        if rank == 0:
            q.put(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpu_tiles_over_nvlink(svo, oracle):
    """Two processes, two GPUs: each renders its interleaved bands; rank 1 stores straight into rank 0's planes through a
    CUDA-IPC mapping (NVLink), one NCCL fence; rank 0's frame equals the oracle's."""
    import ctypes as C
    import socket
    import torch.multiprocessing as mp
    n = C.c_int()
    svo._lib.lib().svo_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    hm, mm = svo.terrain_inputs(128)
    nodes = svo.build_terrain(hm, mm, 128, 64)
    pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
    want, _ = oracle.render(nodes, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=2, max_depth=7), 320, 200, nthreads=4)
    for tag in ("separate", "fused"):
        rgba, depth = out[tag]
        assert np.array_equal(rgba, want["rgba8"]) and np.array_equal(depth.view(np.uint32), want["depth"].view(np.uint32)), tag
    want0, _ = oracle.render(nodes, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=0, max_depth=7), 320, 200, nthreads=4)
    rgba, depth = out["sharedbeam"]
    assert out["beam_planes_equal"], "shared lattice: beam plane differs from svo_beam_conservative's"
    assert np.array_equal(rgba, want0["rgba8"]) and np.array_equal(depth.view(np.uint32), want0["depth"].view(np.uint32)), "sharedbeam"


@pytest.mark.parametrize("kernel", [0, 4, 6])
@pytest.mark.parametrize("cam", ["A", "B", "C"])
@pytest.mark.parametrize("mode", [0, 2, 3, 4])
def test_content_bounds_fast_path_keeps_outputs(svo, oracle, terrain512, cam, mode, kernel):
    """Production configuration (no validation planes): casts that are outside the bounding box of the octree's
    non-empty leaves are ended early (SVO_OPT_CONTENT_BOUNDS).  The reference's outputs -- colour and depth -- must
    not change by a bit; with the option off they must not either."""
    pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
    want, _ = oracle.render(terrain512, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=4, render_mode=mode), 640, 360,
                            nthreads=8, planes=("rgba8", "depth"))
    with svo.SvoContext(640, 360) as c:
        c.set_option(svo._lib.OPT_KERNEL, kernel)  # 4 = upper levels staged in shared memory
        c.upload(terrain512)
        for bounds in (1, 0):
            c.set_option(svo._lib.OPT_CONTENT_BOUNDS, bounds)
            assert c.get_option(svo._lib.OPT_CONTENT_BOUNDS) == bounds
            c.render(svo.camera_frame(cam, frame_number=4, render_mode=mode))
            assert np.array_equal(c.read_color_rgba8(), want["rgba8"]), (cam, mode, bounds)
            d = c.read_depth()
            assert np.array_equal(d.view(np.uint32), want["depth"].view(np.uint32)), (cam, mode, bounds)


def test_content_bounds_inside_camera_and_shallow_depths(svo, oracle, terrain128):
    """Cameras inside the terrain's bounding box, looking up and sideways, and maxDepth values at which interior
    (fill-level) nodes become hits: the box must follow the frame's maxDepth / coneDepth."""
    W, H = 160, 96
    cams = [((1.5, 1.05, 1.5), (-1, 0.2, -1), (-1, 1.5, -1), (1, 0.2, -1), (1, 1.5, -1)),
            ((1.2, 1.12, 1.8), (-1.6, -0.9, -1), (-1.6, 0.9, -1), (1.6, -0.9, -1), (1.6, 0.9, -1)),
            ((1.5, 1.5, 2.0), (-1.6, -0.9, -1), (-1.6, 0.9, -1), (1.6, -0.9, -1), (1.6, 0.9, -1)),
            # far-away origins: t arithmetic loses precision with |origin| (exempt from the shortcut beyond 8)
            ((1.5, 30.0, 1.5), (-0.02, -1, -0.015), (-0.02, -1, 0.015), (0.02, -1, -0.015), (0.02, -1, 0.015)),
            ((7.9, 7.5, 7.7), (-1.05, -1, -0.95), (-1.05, -0.9, -1.0), (-0.95, -1, -1.05), (-0.9, -0.95, -1.0)),
            ((3001.5, 4001.5, 5001.5), (-3.0004, -4.0, -5.0004), (-3.0004, -3.9996, -5.0), (-2.9996, -4.0, -5.0), (-3.0, -4.0004, -4.9996))]
    with svo.SvoContext(W, H) as c:
        c.upload(terrain128)
        for cam in cams:
            for max_depth, cone_depth, mode in ((7, 11, 0), (7, 5, 0), (1, 11, 2), (3, 2, 0), (5, 11, 3)):
                want, _ = oracle.render(terrain128, oracle.make_frame(*cam, frame_number=1, render_mode=mode, max_depth=max_depth,
                                                                      cone_depth=cone_depth), W, H, nthreads=4, planes=("rgba8", "depth"))
                c.render(svo.make_frame(*cam, frame_number=1, render_mode=mode, max_depth=max_depth, cone_depth=cone_depth))
                assert np.array_equal(c.read_color_rgba8(), want["rgba8"]), (cam[0], max_depth, cone_depth, mode)
                assert np.array_equal(c.read_depth().view(np.uint32), want["depth"].view(np.uint32))


def _blob_world(oracle, n=64):
    """BASELINE configs[2] in miniature ("mirrorblobs"): terrain floor + spheres of value 4 (the mirror material)."""
    z, y, x = np.mgrid[0:n, 0:n, 0:n]
    vox = np.zeros((n, n, n), np.uint8)
    h = (6 + 3 * np.sin(x / 7.0) + 2 * np.cos(z / 5.0)).astype(int)
    vox[y <= h] = 1
    vox[(y <= h) & (y >= h - 1)] = 3
    rng = np.random.default_rng(1)  # sphere list from a fixed seed
    for _ in range(6):
        c = rng.uniform(12, n - 12, 3)
        r = rng.uniform(4, 9)
        vox[(x - c[0]) ** 2 + (y - max(c[1], 16)) ** 2 + (z - c[2]) ** 2 <= r * r] = 4
    nodes, _ = oracle.build_dense(vox)
    return nodes


@pytest.mark.parametrize("kernel", [0, 2, 6])
def test_mirror_material_and_deeper_paths(svo, oracle, kernel):
    """svo_frame.casts = 5 (4 bounces) with the mirror rule the shader has commented out (svotrace.comp:500-504) on
    value-4 spheres: the path-traced configuration of BASELINE configs[2], bit-exact against the oracle."""
    nodes = _blob_world(oracle)
    W, H = 192, 108
    with svo.SvoContext(W, H) as c:
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.upload(nodes)
        for cam in ("B", "C"):
            pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
            kw = dict(frame_number=7, render_mode=0, max_depth=6, casts=5, cone_depth=5, mirror_value=4)
            want, st = oracle.render(nodes, oracle.make_frame(pos, l1, l2, r1, r2, **kw), W, H, nthreads=8)
            assert st.casts > 2 * W * H * 0.3  # paths really go deeper than one bounce
            got = _render_gpu(svo, c, svo.camera_frame(cam, **kw))
            _assert_planes_equal(got, want, "mirror cam %s" % cam)


def test_progressive_accumulation(svo, oracle, terrain128):
    """svo_frame.flags bit 0: the running mean over frameNumber the shader has commented out (svotrace.comp:712-719),
    accumulated in the rgba8 framebuffer exactly as written there; frames 1..4 against the oracle."""
    W, H = 160, 90
    pos, l1, l2, r1, r2 = svo.CAMERAS["C"]
    prev = None
    with svo.SvoContext(W, H) as c:
        c.upload(terrain128)
        for frame in (1, 2, 3, 4):
            kw = dict(frame_number=frame, render_mode=0, max_depth=7, flags=1)
            want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, **kw), W, H, nthreads=4,
                                    planes=("rgba8", "depth"), prev_rgba8=prev)
            c.render(svo.camera_frame("C", **kw))
            got = c.read_color_rgba8()
            assert np.array_equal(got, want["rgba8"]), frame
            prev = want["rgba8"]
        first, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=4, render_mode=0, max_depth=7), W, H,
                                 nthreads=4, planes=("rgba8",))
        assert not np.array_equal(first["rgba8"], prev)  # the mean differs from a single sample


def test_ray_binning_keeps_results(svo, oracle, terrain512):
    """SVO_OPT_RAY_SORT: rays are traced in binned order -- 1 = (octant, origin Morton code), 2 = (octant, 64^3 origin cell,
    direction bin, finer origin bits; the default) -- and every hit record lands in the caller's slot, so the output equals the
    unsorted run and the oracle."""
    rng = np.random.default_rng(9)
    n = 100003
    rays = np.zeros(n, dtype=svo.RAY_DTYPE)
    rays["o"] = rng.uniform(1.0, 2.0, (n, 3)).astype(np.float32)
    rays["o"][:, 1] = rng.uniform(1.0, 1.3, n).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["d"][:7] = np.nan
    rays["o"][7:14] = np.nan
    rays["d"][14:21] = 0.0
    rays["d"][21:28] = np.inf
    rays["d"][28:35, 0] = 0.0
    want, _ = oracle.cast_rays(terrain512, rays, max_depth=9, nthreads=8)
    with svo.SvoContext(64, 64) as c:
        c.upload(terrain512)
        assert c.get_option(svo._lib.OPT_RAY_SORT) == 2
        with pytest.raises(svo.SvoError):
            c.set_option(svo._lib.OPT_RAY_SORT, 3)
        for sort in (2, 1, 0):
            c.set_option(svo._lib.OPT_RAY_SORT, sort)
            got = c.cast(rays, max_depth=9)
            for k in ("id", "value", "iter"):
                assert (got[k] == want[k]).all(), (sort, k)
            assert (got["t"].view(np.uint32) == want["t"].view(np.uint32)).all()


def test_pipelined_readback(svo, oracle, terrain128):
    """svo_read_planes_async / svo_swap_buffers: frame s+1 renders while frame s is copied to the host; both frames
    arrive intact."""
    import torch
    W, H = 160, 90
    pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
    with svo.SvoContext(W, H) as c:
        c.upload(terrain128)
        sets = [(torch.empty((H, W, 4), dtype=torch.uint8).pin_memory(), torch.empty((H, W), dtype=torch.float32).pin_memory()) for _ in range(2)]
        for base in (1, 3, 5):
            for k in range(2):
                c.render(svo.camera_frame("B", frame_number=base + k, render_mode=0, max_depth=7))
                c.read_planes_async(sets[k][0].data_ptr(), sets[k][1].data_ptr())
                c.swap_buffers()
            c.read_wait()
            for k in range(2):
                want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=base + k, render_mode=0, max_depth=7),
                                        W, H, nthreads=4, planes=("rgba8", "depth"))
                assert np.array_equal(sets[k][0].numpy(), want["rgba8"]), (base, k)
                assert np.array_equal(sets[k][1].numpy().view(np.uint32), want["depth"].view(np.uint32)), (base, k)
        # a large ray stream in between (its sort scratch is allocated on first use) leaves the frame pipeline intact
        rng = np.random.default_rng(3)
        rays = np.zeros(70000, dtype=svo.RAY_DTYPE)
        rays["o"] = rng.uniform(1.0, 2.0, (rays.size, 3)).astype(np.float32)
        rays["d"] = rng.normal(size=(rays.size, 3)).astype(np.float32)
        got = c.cast(rays, 7)
        want_hits, _ = oracle.cast_rays(terrain128, rays, 7, nthreads=4)
        assert np.array_equal(got["id"], want_hits["id"]) and np.array_equal(got["iter"], want_hits["iter"])
        for k in range(2):
            c.render(svo.camera_frame("B", frame_number=20 + k, render_mode=0, max_depth=7))
            c.read_planes_async(sets[k][0].data_ptr(), sets[k][1].data_ptr())
            c.swap_buffers()
        c.read_wait()
        for k in range(2):
            want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=20 + k, render_mode=0, max_depth=7),
                                    W, H, nthreads=4, planes=("rgba8",))
            assert np.array_equal(sets[k][0].numpy(), want["rgba8"]), k
        # the blocking readers follow the current set
        c.render(svo.camera_frame("B", frame_number=9, render_mode=0, max_depth=7))
        want, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=9, render_mode=0, max_depth=7), W, H,
                                nthreads=4, planes=("rgba8",))
        assert np.array_equal(c.read_color_rgba8(), want["rgba8"])


def test_gpu_transcode_equals_host_transcode(svo, oracle, terrain128, terrain512):
    """svo_upload builds the traversal descriptors on the device (svo_gpu_transcode.cu); the host version
    (svo_transcode.cpp) must give the same descriptors, reference offsets, level count and content bounds."""
    from test_transcode import probe
    import svo_stream as S
    kat = S.serialise(S.interior(1, [S.interior(5, [S.nonsurf(1 if i in (0, 7) else 0) for i in range(8)]), S.surface(2, 955), S.subdiv(0),
                                     S.nonsurf(3)] + [S.nonsurf(0)] * 4))
    streams = [terrain128, terrain512, kat, np.zeros(0, np.uint8), np.zeros(7, np.uint8), np.array([1, 0, 0, 0, 0, 0, 0], np.uint8),
               _blob_world(oracle)]
    with svo.SvoContext(64, 64) as c:
        for nodes in streams:
            want, _ = probe(svo, nodes, 4)
            for gpu in (1, 0):
                c.set_option(svo._lib.OPT_GPU_TRANSCODE, gpu)
                c.upload(nodes)
                assert c.scene_probe() == want, (len(nodes), gpu)
        # partial upload: a scrambled range near the top of the stream (falls back to a whole transcode, or is refused by both paths)
        def canonical_of(nodes):
            with svo.SvoContext(64, 64) as f:
                f.upload(nodes)
                k = f.scene_canonical()
                return k["reachable"], k["hash"], k["depth"]
        edited = terrain128.copy()
        edited[100:200] = terrain128[300:400]
        c.set_option(svo._lib.OPT_GPU_TRANSCODE, 1)
        c.upload(terrain128)
        try:
            c.upload_range(edited, 100, 200)
            k = c.scene_canonical()
            assert (k["reachable"], k["hash"], k["depth"]) == canonical_of(edited)
        except svo.SvoError as e:  # a scrambled stream may legitimately be refused, but then by both paths
            assert e.code == svo._lib.ERR_FORMAT
        # appended bytes (Octree.subdivideNode appends new nodes at memOffset): the range may end beyond the old length
        grown = np.concatenate([terrain128, np.zeros(5000, np.uint8)])
        grown[-7:] = terrain128[:7]
        c.upload(terrain128)
        c.upload_range(grown, terrain128.size, grown.size)
        k = c.scene_canonical()
        assert (k["reachable"], k["hash"], k["depth"]) == canonical_of(grown) and c.scene_info()["stream_bytes"] == grown.size
        assert not c.upload_stats()["whole_transcode"]  # nothing reachable changed


def test_incremental_upload_range_equals_whole_upload(svo, oracle, terrain128):
    """svo_upload_range after edits made the way the engine's SDF brush makes them (Octree.java:700-885), pushed as the two
    byte ranges Renderer.updateSSBO would push (Main.java:349-350): the patched scene has the same descriptor tree as a whole
    upload of the edited stream (layout-independent fingerprint), renders the same frames as the oracle on the edited
    stream -- including where new voxels stick out of the old content box -- and only the touched subtrees were re-walked."""
    import svo_stream as S
    rng = np.random.default_rng(7)
    W, H = 200, 120
    with svo.SvoContext(W, H) as c, svo.SvoContext(64, 64) as fresh:
        c.upload(terrain128)
        stream = terrain128
        for rnd in range(6):
            ed = S.StreamEditor(stream)
            for _ in range(4):
                kind = int(rng.integers(0, 3))
                try:
                    if kind == 0:
                        ed.set_value(ed.find_leaf(1, True, rng), int(rng.integers(0, 4)))
                    elif kind == 1:
                        p = ed.find_leaf(2, bool(rng.integers(0, 2)), rng, min_depth=3)
                        ed.subdivide(p, [int(v) for v in rng.integers(0, 4, 8)])
                        ed.subdivide(p + [int(rng.integers(0, 8))], [int(v) for v in rng.integers(1, 4, 8)], surface=bool(rng.integers(0, 2)))
                    else:
                        ed.set_value(ed.find_leaf(3, True, rng), 0)
                except AssertionError:
                    pass
            new = ed.stream()
            ranges = ed.ranges()
            if rnd % 2:
                ranges = ranges[::-1]
            dirty = 0
            for a, b in ranges:
                c.upload_range(new, a, b)
                st = c.upload_stats()
                assert not st["whole_transcode"], (rnd, st)
                dirty += st["dirty"]
            assert dirty >= 1 and c.scene_info()["stream_bytes"] == new.size
            fresh.upload(new)
            a, b = c.scene_canonical(), fresh.scene_canonical()
            assert (a["reachable"], a["hash"], a["depth"]) == (b["reachable"], b["hash"], b["depth"]), rnd
            for cam, mode in (("B", 0), ("C", 2)):
                pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
                want, _ = oracle.render(new, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=rnd + 1, render_mode=mode, max_depth=7), W, H,
                                        nthreads=8, planes=("rgba8", "depth"))
                c.render(svo.camera_frame(cam, frame_number=rnd + 1, render_mode=mode, max_depth=7))
                assert np.array_equal(c.read_color_rgba8(), want["rgba8"]), (rnd, cam)
                assert np.array_equal(c.read_depth().view(np.uint32), want["depth"].view(np.uint32)), (rnd, cam)
            stream = new
        # pushing the same bytes again changes nothing and costs no transcode
        c.upload_range(stream, 0, min(stream.size, 100000))
        assert c.upload_stats() == {"dirty": 0, "roots": 0, "appended": 0, "whole_transcode": False}


def test_incremental_upload_range_absorbs_the_reference_brush_session(svo, oracle, terrain128):
    """tests/golden/sdf_edits.npz: a session of the REFERENCE'S OWN brush (Octree.useSDFBrush / subdivideNode compiled from its
    Java text, tests/golden/make_sdf_edits.py) -- additive and subtractive spheres, a box, a stroke that changes nothing --
    pushed the way Main.placeSDF pushes it: two svo_upload_range calls per stroke with the ChangeBounds the engine computed.
    After every stroke the scene equals a whole upload of the same bytes and renders the oracle's frames."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdf_edits.npz"))
    assert int(g["n"]) == 128 and int(g["chunk"]) == 64 and int(g["base_bytes"]) == terrain128.size
    W, H = 200, 120
    cur = terrain128
    incremental = 0
    with svo.SvoContext(W, H) as c, svo.SvoContext(64, 64) as fresh:
        c.upload(cur)
        for k in range(int(g["steps"])):
            s0, e0, s1, e1 = (int(v) for v in g["s%d_bounds" % k])
            tail = g["s%d_tail" % k]
            nxt = np.zeros(cur.size + tail.size, np.uint8)
            nxt[:cur.size] = cur
            nxt[g["s%d_idx" % k]] = g["s%d_val" % k]
            nxt[cur.size:] = tail
            if s0 < e0:
                c.upload_range(nxt, s0, e0)    # renderer.updateSSBO(7, buf, cb.start0, cb.end0)
                incremental += not c.upload_stats()["whole_transcode"]
            else:
                with pytest.raises(svo.SvoError):  # "Update SSBO error: Invalid parameters." (Renderer.java:137-140)
                    c.upload_range(nxt, s0, e0)
            if s1 < e1:
                c.upload_range(nxt, s1, e1)    # renderer.updateSSBO(7, buf, cb.start1, cb.end1)
            assert c.scene_info()["stream_bytes"] == nxt.size
            fresh.upload(nxt)
            a, b = c.scene_canonical(), fresh.scene_canonical()
            assert (a["reachable"], a["hash"], a["depth"]) == (b["reachable"], b["hash"], b["depth"]), k
            for cam, mode in (("B", 0), ("C", 2), ("A", 3)):
                pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
                want, _ = oracle.render(nxt, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=k + 1, render_mode=mode, max_depth=7), W, H,
                                        nthreads=8, planes=("rgba8", "depth"))
                c.render(svo.camera_frame(cam, frame_number=k + 1, render_mode=mode, max_depth=7))
                assert np.array_equal(c.read_color_rgba8(), want["rgba8"]), (k, cam)
                assert np.array_equal(c.read_depth().view(np.uint32), want["depth"].view(np.uint32)), (k, cam)
            cur = nxt
    assert incremental >= 3


def test_incremental_upload_range_on_the_bench_world_is_fast(svo):
    """A 1 KB in-place edit of the 8192^3 world (1.95 GB stream, 87 M descriptors): absorbed without a whole transcode, in
    about a millisecond (the whole transcode takes ~0.1 s), with the same descriptor tree as a whole upload."""
    import time
    import svo_stream as S
    size = 8192
    hm, mm = svo.terrain_inputs(size)
    with svo.SvoContext(64, 64) as c:
        c.build_terrain_device(hm, mm, size, 1024)
        nodes = c.download()
        rng = np.random.default_rng(3)
        ed = S.StreamEditor(nodes, slack=4096)
        path = ed.find_leaf(1, True, rng, min_depth=10)
        ed.set_value(path, 0)                      # dig one voxel out of the surface
        _, _, off, _ = ed.walk(path)
        new = ed.stream()
        a = max(0, off - 512)
        c.upload_range(new, a, a + 1024)           # warm-up (allocates the staging buffers)
        assert not c.upload_stats()["whole_transcode"] and c.upload_stats()["dirty"] >= 1
        times = []
        for k in range(5):
            new[off] = (k % 3) + 1
            c.sync()
            t0 = time.perf_counter()
            c.upload_range(new, a, a + 1024)
            c.sync()
            times.append(time.perf_counter() - t0)
            st = c.upload_stats()
            assert not st["whole_transcode"] and st["roots"] == 1, st
        t0 = time.perf_counter()
        c.upload(new)
        c.sync()
        whole = time.perf_counter() - t0
        print("8192^3: 1 KB svo_upload_range %.3f ms (best of 5: %s), whole svo_upload %.1f ms" % (
            1e3 * min(times), ", ".join("%.3f" % (1e3 * t) for t in times), 1e3 * whole))
        assert min(times) < 0.003


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_random_worlds_random_cameras(svo, oracle, seed):
    """Fuzz: random sparse 32^3 / 64^3 voxel worlds (all four record types, deep empty regions, big solid blocks),
    random camera poses inside and outside the cube, random frame parameters -- every plane bit-exact, with the
    validation planes (no shortcuts) and without them (content-bounds shortcut on)."""
    rng = np.random.default_rng(seed)
    n = 32 if seed % 2 else 64
    vox = np.zeros((n, n, n), np.uint8)
    pts = rng.integers(0, n, size=(rng.integers(50, 600), 3))
    vox[pts[:, 2], pts[:, 1], pts[:, 0]] = rng.integers(1, 5, size=len(pts))
    for _ in range(3):
        lo = rng.integers(0, n - 10, 3)
        sz = rng.integers(3, 10, 3)
        vox[lo[2]:lo[2] + sz[2], lo[1]:lo[1] + sz[1], lo[0]:lo[0] + sz[0]] = rng.integers(1, 4)
    nodes, _ = oracle.build_dense(vox)
    W, H = 96, 64
    depth = int(np.log2(n))
    with svo.SvoContext(W, H) as c:
        c.upload(nodes)
        for trial in range(6):
            pos = rng.uniform(0.7, 2.3, 3) if trial % 2 else rng.uniform(1.1, 1.9, 3)
            fwd = rng.normal(size=3)
            fwd /= np.linalg.norm(fwd)
            up = np.cross(fwd, rng.normal(size=3))
            up /= np.linalg.norm(up)
            right = np.cross(fwd, up)
            corners = [fwd + sx * 1.2 * right + sy * 0.8 * up for sx in (-1, 1) for sy in (-1, 1)]
            kw = dict(frame_number=int(rng.integers(1, 50)), render_mode=int(rng.choice([0, 0, 2, 2, 1, 3])),
                      max_depth=int(rng.integers(max(1, depth - 2), depth + 1)), casts=int(rng.integers(1, 4)),
                      cone_depth=int(rng.integers(1, depth + 1)), mirror_value=int(rng.choice([0, 4])))
            of = oracle.make_frame(pos, *corners, **kw)
            want, st = oracle.render(nodes, of, W, H, nthreads=4)
            assert st.stale_pops == 0
            f = svo.make_frame(pos, *corners, **kw)
            c.set_option(svo._lib.OPT_AUX_PLANES, 1)
            _assert_planes_equal(_render_gpu(svo, c, f), want, "seed %d trial %d %s" % (seed, trial, kw))
            c.set_option(svo._lib.OPT_AUX_PLANES, 0)
            c.render(f)
            assert np.array_equal(c.read_color_rgba8(), want["rgba8"]), (seed, trial, kw)
            d = c.read_depth()
            same = (d.view(np.uint32) == want["depth"].view(np.uint32)) | (np.isnan(d) & np.isnan(want["depth"]))
            assert same.all(), (seed, trial, kw)


import os as _os

FULL_SIZES = [2048, 8192]  # BASELINE configs[1] and the bench world


@pytest.mark.parametrize("size", FULL_SIZES)
def test_full_size_properties(svo, oracle, size):
    """BASELINE configs[1] (2048^3) and the bench world (8192^3) at 1920x1080: the FULL mode-0 frame (primary + diffuse
    bounce, colour and depth of all 2 073 600 pixels) against the oracle bit for bit, plus size-independent properties: hit
    ids are offsets of non-empty records of the stream, all kernel variants and the band partition produce the same
    frame, the content-bounds shortcut changes nothing, rendering is deterministic, and a random sample of primary rays
    agrees with the oracle in id, iteration count and t."""
    W, H, depth = 1920, 1080, min(13, int(np.log2(size)))
    hm, mm = svo.terrain_inputs(size)
    nodes = svo.build_terrain(hm, mm, size, 1024)  # the product's generator (byte-equal to the oracle's, test_builder.py)
    rng = np.random.default_rng(5)
    with svo.SvoContext(W, H) as c:
        c.upload(nodes)
        for cam in ("B", "C"):
            f3 = svo.camera_frame(cam, frame_number=1, render_mode=3, max_depth=depth)
            f0 = svo.camera_frame(cam, frame_number=2, render_mode=0, max_depth=depth)
            c.set_option(svo._lib.OPT_KERNEL, 0)
            c.set_option(svo._lib.OPT_AUX_PLANES, 1)
            c.render(f3)
            ids, it, t, dep = c.read_hit_id(), c.read_iter(), c.read_primary_t(), c.read_depth()
            hit = ids != svo.NO_HIT
            assert 0.2 < hit.mean() <= 1.0
            assert (ids[hit] < nodes.size).all() and (nodes[ids[hit]] != 0).all()   # ids name non-empty records
            assert np.isfinite(t[hit]).all() and (t[hit] >= 0).all() and (t[hit] < 4).all()
            assert np.array_equal(dep[hit].view(np.uint32), t[hit].view(np.uint32)) and (dep[~hit] == 0).all()  # mode 3: depth = res.t
            assert it.min() >= 1 and it.max() <= 1501
            # a random sample of primary rays against the oracle (hit id, iteration count, t)
            ys, xs = rng.integers(0, H, 3000), rng.integers(0, W, 3000)
            pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
            fx, fy = ((xs + 0.5) / W).astype(np.float32), ((ys + 0.5) / H).astype(np.float32)
            mix = lambda a, b, w: (np.float32(a) * (np.float32(1) - w) + np.float32(b) * w).astype(np.float32)
            d = np.stack([mix(mix(l1[k], l2[k], fy), mix(r1[k], r2[k], fy), fx) for k in range(3)], 1).astype(np.float32)
            ln = np.sqrt(((d[:, 0] * d[:, 0]) + (d[:, 1] * d[:, 1])) + (d[:, 2] * d[:, 2])).astype(np.float32)
            rays = np.zeros(3000, dtype=svo.RAY_DTYPE)
            rays["o"] = np.asarray(pos, np.float32)
            rays["d"] = (d / ln[:, None]).astype(np.float32)
            want, _ = oracle.cast_rays(nodes, rays, max_depth=depth, nthreads=8)
            assert np.array_equal(ids[ys, xs], want["id"]) and np.array_equal(it[ys, xs], want["iter"])
            assert np.array_equal(t[ys, xs].view(np.uint32), want["t"].view(np.uint32))
            # the frame every variant / partition / option must reproduce
            c.set_option(svo._lib.OPT_AUX_PLANES, 0)
            c.render(f0)
            ref_rgba, ref_depth = c.read_color_rgba8(), c.read_depth()
            c.render(f0)
            assert np.array_equal(c.read_color_rgba8(), ref_rgba)  # deterministic
            full, _ = oracle.render(nodes, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=2, render_mode=0, max_depth=depth), W, H,
                                    nthreads=_os.cpu_count() or 8, planes=("rgba8", "depth"))
            assert np.array_equal(ref_rgba, full["rgba8"]), "full frame colour differs from the oracle (%d^3, cam %s)" % (size, cam)
            assert np.array_equal(ref_depth.view(np.uint32), full["depth"].view(np.uint32)), "full frame depth differs (%d^3, cam %s)" % (size, cam)
            for kernel in (10, 2, 6, 1):
                c.set_option(svo._lib.OPT_KERNEL, kernel)
                c.render(f0)
                assert np.array_equal(c.read_color_rgba8(), ref_rgba), (cam, kernel)
                assert np.array_equal(c.read_depth().view(np.uint32), ref_depth.view(np.uint32)), (cam, kernel)
            c.set_option(svo._lib.OPT_KERNEL, 0)
            c.set_option(svo._lib.OPT_CONTENT_BOUNDS, 0)
            c.render(f0)
            assert np.array_equal(c.read_color_rgba8(), ref_rgba) and np.array_equal(c.read_depth().view(np.uint32), ref_depth.view(np.uint32))
            c.set_option(svo._lib.OPT_CONTENT_BOUNDS, 1)
            c.render(svo.camera_frame("A", frame_number=1, render_mode=3, max_depth=depth))  # scribble
            for part in range(8):
                c.render_interleaved(f0, part, 8)
            assert np.array_equal(c.read_color_rgba8(), ref_rgba) and np.array_equal(c.read_depth().view(np.uint32), ref_depth.view(np.uint32))
            # counters of the instrumented kernel: casts = pixels + primary hits (a bounce is cast iff the primary hit)
            st = c.render_stats(f0)
            assert st["casts"] == W * H + int(hit.sum()) and st["iters"] >= st["casts"]


@pytest.mark.parametrize("kernel", KERNELS)
def test_iteration_cap_boundary(svo, oracle, kernel):
    """Casts that end exactly at, just below and beyond the 1500-iteration cap (svotrace.comp:264-266): the device
    tests the cap only on the POP path, so the boundary needs its own case.  Ray stream + iteration heat map."""
    import svo_stream as S
    nodes = S.serialise(S.tube(10))
    rays = S.tube_rays(10, 100000)
    want, _ = oracle.cast_rays(nodes, rays, 13, nthreads=4)
    assert (want["iter"] == 1500).sum() > 0 and (want["iter"] == 1501).sum() > 0
    W, H = 96, 64
    h = 2.0 ** -10
    cam = ((1.9, 1 + h / 2, 1 + h / 2), (-1, -1e-4, -1e-4), (-1, 1e-4, -1e-4), (-1, -1e-4, 1e-4), (-1, 1e-4, 1e-4))
    with svo.SvoContext(W, H) as c:
        c.upload(nodes)
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        for sort in (0, 1):
            c.set_option(svo._lib.OPT_RAY_SORT, sort)
            got = c.cast(rays, 13)
            for k in ("id", "iter", "value"):
                assert np.array_equal(got[k], want[k]), (k, sort)
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        seen = set()
        for x0 in (1.9, 1.732, 1.7305, 1.73):  # every pixel capped / capped / exactly 1500 iterations / 1497
            pos = (x0,) + cam[0][1:]
            for mode in (1, 0):
                wantp, _ = oracle.render(nodes, oracle.make_frame(pos, *cam[1:], frame_number=3, render_mode=mode, max_depth=13), W, H, nthreads=4)
                c.render(svo.make_frame(pos, *cam[1:], frame_number=3, render_mode=mode, max_depth=13))
                assert np.array_equal(c.read_iter(), wantp["iter"]), (x0, mode)
                assert np.array_equal(c.read_color_rgba8(), wantp["rgba8"]), (x0, mode)
                assert np.array_equal(c.read_radiance().view(np.uint32), wantp["radiance"].view(np.uint32)), (x0, mode)
                seen |= set(np.unique(wantp["iter"]).tolist())
        assert {1497, 1500, 1501} <= seen
