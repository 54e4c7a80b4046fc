"""THE PIN of the builder and of the edit path: the reference's own Java (Octree.java, OctreeThread.java, Util.java,
sdf/*.java), rewritten mechanically and compiled for the CPU by oracle/build_ref_java.py (`_refj` below), against

  * the oracle's restatement of the builder (oracle/svo_builder.c) -- byte for byte, on random dense volumes, shaped
    volumes and heightmap worlds with chunk splices;
  * the product's host builder (csrc/svo_build.cpp: pyramids instead of dense voxels) with the shipped constants
    (CHUNK_SIZE 1024, 512^3 OctreeThreads, maxLOD 9) -- byte for byte on a 1024^3 world (the device builder is byte-compared
    with the host builder in tests/test_gpu_build.py and, emulated, in tests/test_builder.py);
  * the incremental transcode behind svo_upload_range on the SIMT emulator, fed with the edits the engine's brush really
    makes (Octree.useSDFBrush / subdivideNode, Main.placeSDF) and the two ranges its ChangeBounds report.

Needs /root/reference (CPU suite in the build container); tests/golden/sdf_edits.npz carries a recorded session to the GPU box."""
import os

import numpy as np
import pytest

import svo_stream as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdf_edits.npz")


@pytest.fixture(scope="module")
def refj():
    from oracle import ref_java as RJ
    if not RJ.available():
        pytest.skip("no /root/reference and no prebuilt oracle/_ref/libsvo_ref_java.so")
    return RJ


def _shapes(n, rng):
    z, y, x = np.mgrid[0:n, 0:n, 0:n]
    ball = ((x - n / 2) ** 2 + (y - n / 2) ** 2 + (z - n / 2) ** 2 < (n / 3) ** 2).astype(np.uint8) * 2
    slab = (y < n // 2).astype(np.uint8)
    shell = ball.copy()
    shell[((x - n / 2) ** 2 + (y - n / 2) ** 2 + (z - n / 2) ** 2 < (n / 4) ** 2)] = 0
    stripes = ((x // 2 + z) % 3).astype(np.uint8)
    corner = np.zeros((n, n, n), np.uint8)
    corner[: n // 2 + 1, : n // 2, : n // 2 - 1] = 3
    full = np.full((n, n, n), 1, np.uint8)
    empty = np.zeros((n, n, n), np.uint8)
    one = empty.copy()
    one[rng.integers(0, n), rng.integers(0, n), rng.integers(0, n)] = 2
    hole = full.copy()
    hole[rng.integers(0, n), rng.integers(0, n), rng.integers(0, n)] = 0
    return {"ball": ball, "slab": slab, "shell": shell, "stripes": stripes, "corner": corner, "full": full, "empty": empty, "one": one, "hole": hole}


def test_oracle_builder_equals_the_reference_builder_on_dense_volumes(refj, oracle):
    rng = np.random.default_rng(0)
    for n in (2, 4, 8, 16, 32):
        for dens in (0.02, 0.3, 0.7, 0.98):
            v = ((rng.random((n, n, n)) < dens) * rng.integers(1, 4, (n, n, n))).astype(np.uint8)
            a, ca = oracle.build_dense(v)
            b, cb = refj.build_dense(v)
            assert np.array_equal(a, b) and ca == cb, (n, dens, a.size, b.size)
    for n in (8, 32, 64):
        for name, v in _shapes(n, rng).items():
            a, ca = oracle.build_dense(v)
            b, cb = refj.build_dense(v)
            assert np.array_equal(a, b) and ca == cb, (n, name, a.size, b.size)


@pytest.mark.parametrize("n,chunk", [(64, 64), (64, 32), (64, 8), (128, 32), (128, 64), (256, 64), (256, 256)])
def test_all_three_builders_agree_on_heightmap_worlds(refj, oracle, svo, n, chunk):
    hm, mm = svo.terrain_inputs(n)
    a, ca = oracle.build_terrain(hm, mm, n, chunk)
    b, cb = refj.build_terrain(hm, mm, n, chunk)
    assert np.array_equal(a, b) and ca == cb, (n, chunk, a.size, b.size)
    assert np.array_equal(svo.build_terrain(hm, mm, n, chunk), b)
    rng = np.random.default_rng(n + chunk)  # adversarial maps: noise heights, every material
    hm2 = rng.integers(0, 65536, (n, n)).astype(np.uint16)
    mm2 = rng.integers(1, 4, (n, n)).astype(np.uint8)
    b2, _ = refj.build_terrain(hm2, mm2, n, chunk, cap=200 * n * n)
    a2, _ = oracle.build_terrain(hm2, mm2, n, chunk, cap=200 * n * n)
    assert np.array_equal(a2, b2)
    assert np.array_equal(svo.build_terrain(hm2, mm2, n, chunk), b2)


def test_reference_builder_as_shipped_equals_the_product_builder(refj, svo):
    """CHUNK_SIZE 1024, OctreeThreads over 512^3 with maxLOD 9: the constants of Octree.java / OctreeThread.java untouched."""
    n = 1024
    hm, mm = svo.terrain_inputs(n)
    b, _ = refj.build_terrain(hm, mm, n, 1024, cap=64 << 20)
    a = svo.build_terrain(hm, mm, n, 1024)
    assert a.size == b.size and np.array_equal(a, b)


def _replay(cur, step, g):
    """The stream after a recorded stroke as the GPU copy sees it (only the pushed ranges change), and the ranges."""
    s0, e0, s1, e1 = (int(v) for v in g["s%d_bounds" % step])
    new_size = cur.size + g["s%d_tail" % step].size
    nxt = np.zeros(new_size, np.uint8)
    nxt[:cur.size] = cur
    nxt[g["s%d_idx" % step]] = g["s%d_val" % step]
    nxt[cur.size:] = g["s%d_tail" % step]
    ranges = [(a, b) for a, b in ((s0, e0), (s1, e1)) if b > a]
    return nxt, ranges


def test_recorded_session_is_what_the_reference_brush_does(refj, oracle, svo):
    """tests/golden/sdf_edits.npz replays to the streams the compiled brush produces now (the fixture is not stale), the
    bytes the engine changes outside its own ChangeBounds are DELETE_VALUE marks on unreachable records, and every stroke
    changes voxels inside its brush's box only."""
    g = np.load(GOLDEN)
    n, chunk, lod = int(g["n"]), int(g["chunk"]), int(g["max_lod"])
    hm, mm = svo.terrain_inputs(n)
    cur, _ = oracle.build_terrain(hm, mm, n, chunk)
    assert cur.size == int(g["base_bytes"])
    engine = cur.copy()  # the engine's own (host) copy
    stale_all = set()
    for k in range(int(g["steps"])):
        what = g["s%d_what" % k]
        kind, value, origin, params = ("sphere", "box")[int(what[0])], int(what[1]), tuple(int(v) for v in what[2:5]), tuple(int(v) for v in what[5:])
        new, ranges, cb = refj.sdf_brush(engine, n, lod, origin, params, value, kind=kind)
        gpu, granges = _replay(cur, k, g)
        assert granges == ranges and tuple(int(v) for v in g["s%d_bounds" % k]) == cb
        stale_all |= set(int(i) for i in g["s%d_stale" % k])  # the GPU copy misses these writes until a later range covers them
        differ = np.nonzero(gpu != new)[0]
        assert set(int(i) for i in differ) <= stale_all and (new[differ] == 127).all()
        before, after = S.decode_voxels(engine, n), S.decode_voxels(new, n)
        zz, yy, xx = np.nonzero(before != after)
        if zz.size:
            r = (2 if kind == "box" else 1) * max(params) + 2  # (Box.distance takes width/height/depth as HALF extents, Box.java:37-51)
            assert (abs(xx - origin[0]) <= r).all() and (abs(yy - origin[1]) <= r).all() and (abs(zz - origin[2]) <= r).all()
            assert set(np.unique(after[zz, yy, xx])) <= {value}
        # the GPU copy (stale marks missing) decodes to the same voxels: the marks sit on records nothing points to any more
        assert np.array_equal(S.decode_voxels(gpu, n), after)
        engine, cur = new, gpu


def test_incremental_transcode_absorbs_the_reference_brush_session():
    """Every recorded stroke pushed as the engine pushes it (two ranges) through gpu_diff_apply + gpu_patch on the SIMT
    emulator: the patched descriptor tree equals a whole transcode of the same bytes."""
    from hostemu import emu as E
    import svo_raytracer_b200 as svo
    from oracle import oracle as O
    g = np.load(GOLDEN)
    n, chunk = int(g["n"]), int(g["chunk"])
    hm, mm = svo.terrain_inputs(n)
    cur, _ = O.build_terrain(hm, mm, n, chunk)
    patched_steps = 0
    for k in range(int(g["steps"])):
        nxt, ranges = _replay(cur, k, g)
        if ranges:
            r = E.patch_check(cur, nxt, ranges)
            assert r["status"] == 0, (k, r)
            patched_steps += r["fell_back"] == 0
            assert r["fell_back"] == 1 or r["stored"] >= r["reachable"] > 0, (k, r)
        else:
            assert np.array_equal(cur, nxt)
        cur = nxt
    assert patched_steps >= 4


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_brush_sessions_through_the_incremental_transcode(refj, oracle, svo, seed):
    """Random strokes of the reference's brush (spheres and boxes, additive and subtractive, on / under / above the surface),
    each pushed as the engine pushes it: the GPU copy is the old bytes with the two ranges overwritten, the patched
    descriptor tree equals a whole transcode of those bytes (SIMT emulator), and it decodes to the engine's own voxels."""
    from hostemu import emu as E
    n = 128
    hm, mm = svo.terrain_inputs(n)
    base, _ = oracle.build_terrain(hm, mm, n, 64)
    hs = (hm.astype(np.uint32) * (n // 4)) >> 16
    rng = np.random.default_rng(seed)
    engine, gpu = base.copy(), base.copy()
    incremental = 0
    for k in range(6):
        x, z = int(rng.integers(4, n - 4)), int(rng.integers(4, n - 4))
        y = int(hs[z, x]) + int(rng.integers(-6, 10))
        kind = "box" if rng.random() < 0.25 else "sphere"
        params = tuple(int(v) for v in rng.integers(1, 7, 3)) if kind == "box" else (int(rng.integers(1, 10)),)
        value = int(rng.integers(0, 4))
        new, ranges, _ = refj.sdf_brush(engine, n, 7, (x, y, z), params, value, kind=kind)
        nxt = np.zeros(new.size, np.uint8)
        nxt[:gpu.size] = gpu
        for a, b in ranges:
            nxt[a:b] = new[a:b]
        if ranges:
            r = E.patch_check(gpu, nxt, ranges)
            assert r["status"] == 0, (seed, k, kind, value, (x, y, z), params, r)
            incremental += r["fell_back"] == 0
        else:
            assert np.array_equal(nxt, gpu)
        assert np.array_equal(S.decode_voxels(nxt, n), S.decode_voxels(new, n)), (seed, k)
        engine, gpu = new, nxt
    assert incremental >= 3
