"""THE PIN.  oracle/_ref/libsvo_ref.so is the reference's own svotrace.comp / svobeam.comp compiled for the CPU
(oracle/build_ref.py reads the shader text from /root/reference and compiles it through oracle/glsl_shim.h).  These
tests demand that oracle/svo_oracle.c -- the restatement every GPU parity test is measured against -- equals it BIT
FOR BIT: every output plane of every render mode on BASELINE configs[0], single casts with every result field, ray
streams, the 1500-iteration cap boundary, zero / axis-parallel / NaN directions, stale in/out fields, the
generalised constants (maxDepth, cone depth, cast count) and the beam pre-pass.

Where /root/reference is absent (the GPU box) the prebuilt library travels with the snapshot; with neither, the
tests skip -- the committed fixtures under tests/golden/ (generated FROM this library) carry the pin there."""
import numpy as np
import pytest

PLANES = ("rgba8", "depth", "radiance", "hit_id", "iter", "primary_t")


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as R
    if not R.available():
        pytest.skip("no /root/reference to compile and no prebuilt oracle/_ref/libsvo_ref.so")
    R.lib()
    return R


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    return a == b


def _assert_planes(got, want, what):
    for k in PLANES:
        bad = int((~_same(got[k], want[k])).sum())
        assert bad == 0, "%s: plane %s differs in %d of %d elements" % (what, k, bad, got[k].size)


def _fields(res):
    out = [res.value, res.pointer, res.iter, res.depth]
    f = [res.t, res.scale] + list(res.hitPos) + list(res.debugColor) + list(res.normal) + list(res.voxelPos)
    return out, np.array(f, np.float32)


def _compare_casts(oracle, ref, nodes, rays, max_depth, cone, cone_depth=11, stale=False):
    hits = capped = 0
    for o, d in rays:
        ra, rb = oracle.CastResult(), oracle.CastResult()
        if stale:  # in/out parameter: fields the call does not write must survive (U4)
            for r in (ra, rb):
                r.value, r.pointer, r.iter, r.t, r.scale, r.depth = 9, 77, 1234, 0.375, 0.125, 5
                r.normal[:] = [0.5, -0.25, 1.0]
                r.voxelPos[:] = [1.25, 1.5, 1.75]
                r.hitPos[:] = [3.0, 4.0, 5.0]
        ha, ra, st = oracle.cast(nodes, o, d, max_depth=max_depth, cone_trace=cone, cone_depth=cone_depth, res=ra)
        hb, rb, loop_iter = ref.cast(nodes, o, d, max_depth=max_depth, cone_trace=cone, cone_depth=cone_depth, res=rb)
        what = (tuple(float(v) for v in o), tuple(float(v) for v in d), max_depth, cone, cone_depth)
        assert ha == hb, what
        # the oracle's Stats.iters counts iterations that fetched a child: 1500 for a cast the cap ended at its 1501st
        assert st.iters == min(loop_iter, 1500) and st.capped == (loop_iter > 1500), what
        ia, fa = _fields(ra)
        ib, fb = _fields(rb)
        assert ia == ib, what + (ia, ib)
        assert _same(fa, fb).all(), what + (fa, fb)
        hits += ha
        capped += loop_iter > 1500
    return hits, capped


def _rays(rng, n):
    out = []
    for i in range(n):
        o = rng.uniform(0.8, 2.2, 3) if i % 3 else rng.uniform(1.05, 1.95, 3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        if i % 41 == 0:
            d[rng.integers(0, 3)] = 0.0   # axis-parallel: EPSILON * sign(0) = 0 -> coefficient -inf
        if i % 43 == 0:
            d[rng.integers(0, 3)] = -0.0
        if i % 47 == 0:
            d[rng.integers(0, 3)] = 1e-20  # below EPSILON: clamped to +-2^-48
        if i % 97 == 0:
            d[:] = 0.0
        if i % 101 == 0:
            d[rng.integers(0, 3)] = np.nan
        if i % 103 == 0:
            d[rng.integers(0, 3)] = np.inf
        out.append((o.astype(np.float32), d.astype(np.float32)))
    return out


@pytest.mark.parametrize("cam", ["A", "B", "C"])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_config0_every_mode_and_camera(svo, oracle, ref, terrain512, cam, mode):
    """BASELINE configs[0] (512^3 terrain, 640x360) as shipped: MAX_DEPTH 13, 2 casts, cone depth 11."""
    pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
    f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=3, render_mode=mode)
    want = ref.render(terrain512, f, 640, 360, nthreads=8)
    got, _ = oracle.render(terrain512, f, 640, 360, nthreads=8)
    _assert_planes(got, want, "cam %s mode %d" % (cam, mode))
    if mode != 4:
        assert (want["hit_id"] != 0xFFFFFFFF).mean() > 0.1


@pytest.mark.parametrize("kw", [dict(max_depth=9), dict(max_depth=7), dict(max_depth=8, cone_depth=6),
                                dict(max_depth=9, casts=1), dict(max_depth=9, casts=4, cone_depth=8),
                                dict(max_depth=13, frame_number=99), dict(max_depth=9, frame_number=1000)],
                         ids=lambda kw: "-".join("%s%d" % (k[:4], v) for k, v in kw.items()))
def test_generalised_constants(svo, oracle, ref, terrain128, terrain512, kw):
    """The oracle turns MAX_DEPTH, the cone cut's 11 and the mode-0 loop bound 2 into frame parameters; build_ref.py
    substitutes the same three constants by variables in the shader text (R9).  Interior hits at maxDepth (the
    child-type mask decoded as a normal) are on this path."""
    for nodes, (w, h) in ((terrain512, (320, 180)), (terrain128, (161, 91))):
        for cam in ("B", "C"):
            pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
            for mode in (0, 2):
                f = oracle.make_frame(pos, l1, l2, r1, r2, render_mode=mode, **kw)
                want = ref.render(nodes, f, w, h, nthreads=8)
                got, _ = oracle.render(nodes, f, w, h, nthreads=8)
                _assert_planes(got, want, "%s cam %s mode %d %dx%d" % (kw, cam, mode, w, h))


def test_row_bands(svo, oracle, ref, terrain128):
    """Rows [y0, y1) only: what the multi-GPU tile partition renders per rank."""
    pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
    f = oracle.make_frame(pos, l1, l2, r1, r2, render_mode=0, max_depth=7)
    want = ref.render(terrain128, f, 200, 120, y0=37, y1=90, nthreads=4)
    got, _ = oracle.render(terrain128, f, 200, 120, y0=37, y1=90, nthreads=4)
    _assert_planes(got, want, "rows 37..90")
    assert want["rgba8"][:37].max() == 0 and want["rgba8"][90:].max() == 0 and want["rgba8"][37:90].max() > 0


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_worlds_single_casts(oracle, ref, seed):
    rng = np.random.default_rng(seed)
    n = 16 if seed == 1 else 32
    vox = np.zeros((n, n, n), np.uint8)
    pts = rng.integers(0, n, size=(rng.integers(30, 400), 3))
    vox[pts[:, 2], pts[:, 1], pts[:, 0]] = rng.integers(1, 5, size=len(pts))
    lo = rng.integers(0, n - 8, 3)
    vox[lo[2]:lo[2] + 7, lo[1]:lo[1] + 5, lo[0]:lo[0] + 6] = 2  # a solid block: interior + non-surface leaves
    nodes, _ = oracle.build_dense(vox)
    depth = int(np.log2(n))
    hits = 0
    for max_depth, cone, cd in ((depth, False, 11), (depth - 1, False, 11), (depth, True, 11), (depth, True, 3), (2, False, 11), (13, False, 11)):
        h, _ = _compare_casts(oracle, ref, nodes, _rays(rng, 250), max_depth, cone, cd, stale=(max_depth == depth))
        hits += h
    assert hits > 100


def test_terrain_single_casts(oracle, ref, terrain128):
    rng = np.random.default_rng(5)
    rays = _rays(rng, 300)
    for _ in range(150):  # rays that graze the surface from above
        o = np.array([rng.uniform(1.0, 2.0), rng.uniform(1.2, 1.4), rng.uniform(1.0, 2.0)], np.float32)
        d = np.array([rng.normal(), -abs(rng.normal()) * 0.2, rng.normal()])
        rays.append((o, (d / np.linalg.norm(d)).astype(np.float32)))
    hits, _ = _compare_casts(oracle, ref, terrain128, rays, 7, False)
    assert hits > 50
    _compare_casts(oracle, ref, terrain128, rays[:200], 7, True, stale=True)
    _compare_casts(oracle, ref, terrain128, rays[:200], 5, False)


def test_hand_assembled_streams_and_iteration_cap(oracle, ref):
    import svo_stream as S
    # all four record types under one root, packed normals including the NaN one (555)
    kids = [S.surface(1, normal=0), S.nonsurf(2), S.subdiv(3), S.surface(2, normal=555), S.nonsurf(0), S.surface(3, normal=999),
            S.interior(1, [S.surface(1, normal=123), S.nonsurf(0), S.nonsurf(1), S.subdiv(0), S.surface(2, normal=987), S.nonsurf(0),
                           S.nonsurf(0), S.surface(3, normal=505)]), S.subdiv(0)]
    nodes = S.serialise(S.interior(1, kids))
    rng = np.random.default_rng(9)
    hits, _ = _compare_casts(oracle, ref, nodes, _rays(rng, 400), 13, False)
    assert hits > 50
    _compare_casts(oracle, ref, nodes, _rays(rng, 200), 1, False, stale=True)
    # truncated stream: records that run past the end of the buffer read zeros (U1)
    _compare_casts(oracle, ref, nodes[:len(nodes) - 9], _rays(rng, 200), 13, False)
    # the tube world: casts that end just below, at and beyond the 1500-iteration cap
    tube = S.serialise(S.tube(10))
    rays = S.tube_rays(10, 4000)
    want = ref.cast_rays(tube, rays, 13, nthreads=4)
    got, _ = oracle.cast_rays(tube, rays, 13, nthreads=4)
    assert want.tobytes() == got.tobytes()
    assert (want["iter"] == 1501).any() and (want["iter"] == 1500).any() and ((want["iter"] > 1400) & (want["iter"] < 1500)).any()
    pick = np.concatenate([np.flatnonzero(want["iter"] == k)[:3] for k in (1501, 1500, 1499, 1497)] + [np.arange(5)])
    _, capped = _compare_casts(oracle, ref, tube, [(rays["o"][i], rays["d"][i]) for i in pick], 13, False)
    assert capped >= 1


def test_ray_streams(oracle, ref, terrain128, terrain512):
    """BASELINE configs[3] in miniature: incoherent rays from in and around the cube."""
    rng = np.random.default_rng(11)
    n = 60000
    rays = np.zeros(n, oracle.RAY_DTYPE)
    rays["o"] = rng.uniform(0.7, 2.3, (n, 3))
    d = rng.normal(size=(n, 3))
    rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["d"][::37, 1] = 0.0
    rays["d"][::1013] = np.nan
    for nodes, depth in ((terrain128, 7), (terrain512, 9), (terrain512, 13)):
        want = ref.cast_rays(nodes, rays, depth, nthreads=8)
        got, _ = oracle.cast_rays(nodes, rays, depth, nthreads=8)
        assert want.tobytes() == got.tobytes()
        assert (want["id"] != 0xFFFFFFFF).mean() > 0.1


def _world_without_subdividable_leaves(oracle, n, seed):
    """svobeam.comp's own extractChild has no case for type code 2 (svobeam.comp:124-143): on a tree with
    subdividable leaves control falls off the end of a non-void function.  This world has none: in every aligned
    2x2x2 block two voxels differ, so no region larger than one voxel is homogeneous."""
    rng = np.random.default_rng(seed)
    vox = (rng.random((n, n, n)) < 0.12).astype(np.uint8) * rng.integers(1, 4, (n, n, n)).astype(np.uint8)
    z, y, x = np.mgrid[0:n, 0:n, 0:n]
    vox[y < n // 3] = 1
    vox[(x % 2 == 0) & (y % 2 == 0) & (z % 2 == 0)] = 0
    vox[(x % 2 == 1) & (y % 2 == 0) & (z % 2 == 0)] = 2
    nodes, counts = oracle.build_dense(vox)
    assert counts[2] == 0
    return nodes


@pytest.mark.parametrize("cam", ["A", "B", "C"])
def test_beam_prepass(svo, oracle, ref, cam):
    """svobeam.comp main() compiled unchanged vs svo_oracle_beam, then the fine pass consuming the beam image in both."""
    nodes = _world_without_subdividable_leaves(oracle, 32, 3)
    pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
    w, h = 160, 92
    f = oracle.make_frame(pos, l1, l2, r1, r2, render_mode=2, use_beam=1, max_depth=5)
    want = ref.beam(nodes, f, w, h)
    got = oracle.beam(nodes, f, w, h)
    assert _same(got, want).all(), "beam image differs in %d texels" % int((~_same(got, want)).sum())
    if cam != "B":  # camera B starts inside this world's solid third: every beam distance is 0
        assert (want > 0).mean() > 0.05
    for mode in (0, 2):
        f = oracle.make_frame(pos, l1, l2, r1, r2, render_mode=mode, use_beam=1, max_depth=5)
        a = ref.render(nodes, f, w, h, beam=want, nthreads=4)
        b, _ = oracle.render(nodes, f, w, h, beam=want, nthreads=4)
        _assert_planes(b, a, "fine pass with beam, mode %d" % mode)


def test_ref_refuses_what_the_shader_does_not_have(svo, oracle, ref, terrain128):
    pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
    for kw in (dict(mirror_value=4), dict(flags=1)):
        with pytest.raises(ValueError):
            ref.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, **kw), 32, 16)


def test_shim_arithmetic_is_the_contract(oracle, ref):
    """glsl_shim.h implements the contract with oracle_math.h's own kernels; rand() comes from the shader text."""
    about = ref.lib().svo_ref_about()
    assert b"svotrace.comp" in about
    # rand(): one pixel's mode-0 bounce direction depends on it; covered by the frames above.  Here: the shim's
    # image store rule U5 on the values that matter (NaN, negatives, >1, exact halves).
    import svo_stream as S
    nodes = S.serialise(S.interior(1, [S.nonsurf(0)] * 8))
    f = oracle.make_frame((1.5, 1.5, 3.0), (-1, -1, -1), (-1, 1, -1), (1, -1, -1), (1, 1, -1), render_mode=2)
    a = ref.render(nodes, f, 24, 24)
    b, _ = oracle.render(nodes, f, 24, 24)
    _assert_planes(b, a, "empty world (sky + overlay)")
    assert (a["rgba8"][:10, :10, :3] == 255).all()
    z = np.zeros(7, np.uint8)  # octreeBuffer[0] == 0 -> red overlay (svotrace.comp:696-698)
    a = ref.render(z, f, 16, 16)
    b, _ = oracle.render(z, f, 16, 16)
    _assert_planes(b, a, "zero buffer")
    assert (a["rgba8"][0, 0] == [255, 0, 0, 255]).all()
