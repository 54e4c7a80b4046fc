"""The C oracle against a second restatement of the reference traversal written independently from the GLSL text
(tests/glsl_restatement.py): hit flag, iteration count, hit id, t, value, packed-normal decode, scale / depth and voxelPos
must agree bit for bit on random voxel worlds, on a terrain with chunk splices and on hand-assembled streams, for rays
from inside and outside the cube, axis-parallel, zero and NaN directions, every maxDepth, with and without the cone cut."""
import numpy as np
import pytest

import glsl_restatement as G


def _u32(x):
    return int(np.array([x], dtype=np.float32).view(np.uint32)[0])


def _compare(oracle, nodes, rays, max_depth, cone):
    buf = G.Buffer(nodes)
    hits = capped = 0
    for o, d in rays:
        want_hit, res, st = oracle.cast(nodes, o, d, max_depth=max_depth, cone_trace=cone, cone_depth=11)
        got = G.intersect_octree(buf, o, d, max_depth, cone)
        what = (tuple(float(v) for v in o), tuple(float(v) for v in d), max_depth, cone)
        assert got["stale_pops"] == 0 and st.stale_pops == 0, what
        assert got["hit"] == want_hit, what
        # Stats.iters counts the iterations that fetched a child: 1500 for a cast the cap ended at its 1501st
        assert (st.iters, st.capped) == ((1500, 1) if got["capped"] else (got["iter"], 0)), what
        capped += got["capped"]
        if got["capped"] or "pointer" not in got:
            continue
        hits += want_hit
        if want_hit:  # the fields intersectOctree writes after the loop (:403-428)
            assert got["pointer"] == res.pointer and got["value"] == res.value and got["depth"] == res.depth and got["iter"] == res.iter, what
            assert _u32(got["t"]) == _u32(res.t) and _u32(got["scale"]) == _u32(res.scale), what
            for k in range(3):
                a, b = got["normal"][k], np.float32(res.normal[k])
                assert _u32(a) == _u32(b) or (np.isnan(a) and np.isnan(b)), what
                a, b = got["voxelPos"][k], np.float32(res.voxelPos[k])
                assert _u32(a) == _u32(b) or (np.isnan(a) and np.isnan(b)), what
    return hits, capped


def _rays(rng, n):
    out = []
    for i in range(n):
        o = rng.uniform(0.8, 2.2, 3) if i % 3 else rng.uniform(1.05, 1.95, 3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        if i % 41 == 0:
            d[rng.integers(0, 3)] = 0.0  # axis-parallel: EPSILON * sign(0) = 0 -> coefficient -inf
        if i % 97 == 0:
            d[:] = 0.0
        if i % 101 == 0:
            d[rng.integers(0, 3)] = np.nan
        out.append((o.astype(np.float32), d.astype(np.float32)))
    return out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_worlds(oracle, seed):
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
    for max_depth, cone in ((depth, False), (depth - 1, False), (depth, True), (2, False), (13, False)):
        h, _ = _compare(oracle, nodes, _rays(rng, 250), max_depth, cone)
        hits += h
    assert hits > 100


def test_terrain_with_chunk_splices(oracle, terrain128):
    rng = np.random.default_rng(5)
    rays = _rays(rng, 300)
    for k in range(100):  # rays that graze the surface from above
        o = np.array([rng.uniform(1.0, 2.0), rng.uniform(1.2, 1.4), rng.uniform(1.0, 2.0)], np.float32)
        d = np.array([rng.normal(), -abs(rng.normal()) * 0.2, rng.normal()])
        rays.append((o, (d / np.linalg.norm(d)).astype(np.float32)))
    hits, _ = _compare(oracle, terrain128, rays, 7, False)
    assert hits > 50
    _compare(oracle, terrain128, rays[:150], 7, True)
    _compare(oracle, terrain128, rays[:150], 5, False)


def test_hand_assembled_streams_and_iteration_cap(oracle):
    import svo_stream as S
    # all four record types under one root, packed normals including the NaN one (555)
    kids = [S.surface(1, normal=0), S.nonsurf(2), S.subdiv(3), S.surface(2, normal=555), S.nonsurf(0), S.surface(3, normal=999),
            S.interior(1, [S.surface(1, normal=123), S.nonsurf(0), S.nonsurf(1), S.subdiv(0), S.surface(2, normal=987), S.nonsurf(0),
                           S.nonsurf(0), S.surface(3, normal=505)]), S.subdiv(0)]
    nodes = S.serialise(S.interior(1, kids))
    rng = np.random.default_rng(9)
    hits, _ = _compare(oracle, nodes, _rays(rng, 400), 13, False)
    assert hits > 50
    _compare(oracle, nodes, _rays(rng, 200), 1, False)
    # the tube world: casts that end just below, at and beyond the 1500-iteration cap
    tube = S.serialise(S.tube(10))
    rays = S.tube_rays(10, 4000)
    want, _ = oracle.cast_rays(tube, rays, 13, nthreads=4)
    pick = np.concatenate([np.flatnonzero(want["iter"] == k)[:3] for k in (1501, 1500, 1499, 1497)] + [np.arange(5)])
    assert (want["iter"][pick] == 1501).any() and (want["iter"][pick] == 1500).any()
    _, capped = _compare(oracle, tube, [(rays["o"][i], rays["d"][i]) for i in pick], 13, False)
    assert capped >= 1


def _math(oracle):
    L = oracle.lib()
    return {k: (lambda f: (lambda x: np.float32(f(float(x)))))(getattr(L, "svo_oracle_" + k)) for k in ("sin", "cos", "acos", "exp")}


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_trace_and_main_every_render_mode(svo, oracle, mode):
    """trace() and main() restated a second time (the contract's sin / cos / acos / exp kernels are shared, everything else --
    bounce loop, stale castResult fields, material table, fog, shadow ray, penumbra, debug overlay -- is written anew):
    colour, depth, primary hit id, iteration count and t of every pixel of a small frame, bit for bit."""
    rng = np.random.default_rng(17)
    n = 32
    z, y, x = np.mgrid[0:n, 0:n, 0:n]
    vox = np.zeros((n, n, n), np.uint8)
    h = (9 + 4 * np.sin(x / 5.0) + 3 * np.cos(z / 4.0)).astype(int)
    vox[y <= h] = 1
    vox[(y <= h) & (y >= h - 1)] = 3
    vox[(x - 20) ** 2 + (y - 20) ** 2 + (z - 12) ** 2 <= 30] = 2
    pts = rng.integers(0, n, size=(40, 3))
    vox[pts[:, 2], pts[:, 1], pts[:, 0]] = 4  # a value outside the material table
    nodes, _ = oracle.build_dense(vox)
    buf, m = G.Buffer(nodes), _math(oracle)
    W, H = 40, 24
    for cam in ("B", "C"):
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        kw = dict(frame_number=5, render_mode=mode, max_depth=5, casts=2 if cam == "B" else 3)
        want, _ = oracle.render(nodes, oracle.make_frame(pos, l1, l2, r1, r2, **kw), W, H, nthreads=4)
        for yy in range(H):
            for xx in range(W):
                color, depth, primary = G.render_pixel(buf, m, pos, l1, l2, r1, r2, 5, mode, W, H, xx, yy, max_depth=5, casts=kw["casts"])
                what = (cam, mode, xx, yy)
                for c in range(3):
                    a, b = np.float32(color[c]), want["radiance"][yy, xx, c]
                    assert _u32(a) == _u32(b) or (np.isnan(a) and np.isnan(b)), what + ("colour", c, a, b)
                a, b = np.float32(depth), want["depth"][yy, xx]
                assert _u32(a) == _u32(b) or (np.isnan(a) and np.isnan(b)), what + ("depth", a, b)
                if primary is not None:  # mode 4 casts nothing
                    assert min(primary["iter"], 1501) == want["iter"][yy, xx], what
                    assert (primary["pointer"] if primary["hit"] else 0xFFFFFFFF) == want["hit_id"][yy, xx], what
                    if primary["hit"]:
                        assert _u32(primary["t"]) == _u32(want["primary_t"][yy, xx]), what
