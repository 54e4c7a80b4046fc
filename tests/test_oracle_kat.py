"""Known-answer tests that pin the CPU oracle (oracle/svo_oracle.c).

The reference ships no golden vector for its traversal ("parity
unpinned"), so the oracle is pinned against answers derived by hand or by an
independent float64 geometric computation:
  1 root-only tree, one solid child: pointer = 7 + child, analytic t
  2 all 8 ray octants x all 8 children (mirroring, child_shift = idx ^ octant_mask)
  3 mixed-type sibling block: the byte-offset scan of extractChild (codes 0/1/2/3 -> 7/3/7/1 bytes)
  4 a single voxel at depth 13 (scale stack / mantissa logic down to 2^-13)
  5 axis-parallel ray (dir component 0 -> -inf coefficient, NaN corner; min/max must ignore NaN)
  6 NaN direction: spins to the 1500-iteration cap and misses
  7 packed normals: 955 -> (0,0,1); 555 -> normalize(0) = NaN
  8 random sparse 32^3 worlds: first hit agrees with a brute-force float64 slab test over all solid voxels
"""
import itertools
import math

import numpy as np
import pytest

import svo_stream as S

CUBE_LO, CUBE_HI = 1.0, 2.0


def child_box(k, lo=(1.0, 1.0, 1.0), size=1.0):
    h = size / 2
    c = [lo[a] + ((k >> a) & 1) * h for a in range(3)]
    return c, [c[a] + h for a in range(3)]


def root_with(children):
    return S.serialise(S.interior(1, children))


def aim(box_lo, box_hi, d, back=1.7):
    d = np.asarray(d, np.float64)
    d = d / np.linalg.norm(d)
    c = (np.asarray(box_lo) + np.asarray(box_hi)) / 2
    return (c - back * d).astype(np.float32), d.astype(np.float32)


def test_kat1_root_only_single_child(oracle):
    k = 4
    nodes = root_with([S.nonsurf(5 if i == k else 0) for i in range(8)])
    assert nodes.size == 15
    lo, hi = child_box(k)
    o, d = aim(lo, hi, (0.3, 0.2, -1.0))
    hit, res, st = oracle.cast(nodes, o, d)
    assert hit and res.pointer == 7 + k and res.value == 5
    assert res.depth == 1 and res.scale == 0.5
    assert res.t == pytest.approx(S.slab_hit(o, d, lo, hi), rel=1e-6)
    assert list(res.normal) == [0.0, 0.0, 0.0]
    assert list(res.voxelPos) == pytest.approx(lo)
    assert st.stale_pops == 0 and st.capped == 0


@pytest.mark.parametrize("signs", list(itertools.product((-1, 1), repeat=3)))
def test_kat2_octants_times_children(oracle, signs):
    for k in range(8):
        nodes = root_with([S.nonsurf(3 if i == k else 0) for i in range(8)])
        lo, hi = child_box(k)
        o, d = aim(lo, hi, (signs[0] * 1.0, signs[1] * 0.8, signs[2] * 0.6))
        hit, res, _ = oracle.cast(nodes, o, d)
        assert hit, (signs, k)
        assert res.pointer == 7 + k and res.value == 3
        assert res.t == pytest.approx(S.slab_hit(o, d, lo, hi), rel=1e-6)
        assert list(res.voxelPos) == pytest.approx(lo)
    # nothing solid: miss from every octant, and the miss leaves `res` untouched (stale-field semantics)
    nodes = root_with([S.nonsurf(0) for _ in range(8)])
    o, d = aim((1, 1, 1), (2, 2, 2), signs)
    hit, res, _ = oracle.cast(nodes, o, d)
    assert not hit and res.t == 0.0 and res.value == 0


def test_kat3_mixed_type_sibling_block(oracle):
    kinds = ["interior", "surface", "subdiv", "nonsurf", "surface", "nonsurf", "subdiv", "interior"]
    sizes = {"interior": 7, "surface": 3, "subdiv": 7, "nonsurf": 1}
    for k in range(8):
        ch = []
        for i, kind in enumerate(kinds):
            v = 9 if i == k else 0
            if kind == "interior":
                ch.append(S.Node("interior", v, None))  # code 0 with cp == 0: a leaf to the traversal (svotrace.comp:311)
            elif kind == "surface":
                ch.append(S.surface(v, 955))
            elif kind == "subdiv":
                ch.append(S.subdiv(v))
            else:
                ch.append(S.nonsurf(v))
        nodes = root_with(ch)
        assert nodes.size == 7 + sum(sizes[x] for x in kinds)
        lo, hi = child_box(k)
        o, d = aim(lo, hi, (-0.5, -0.7, -1.0))
        hit, res, _ = oracle.cast(nodes, o, d)
        assert hit and res.value == 9
        assert res.pointer == 7 + sum(sizes[x] for x in kinds[:k]), k
        if kinds[k] == "surface":  # 955 -> digits (5,5,9) - 5 = (0,0,4) -> (0,0,1)
            assert list(res.normal) == [0.0, 0.0, 1.0]
            assert res.voxelPos[2] == pytest.approx(lo[2] + 1.0 * 0.5 * 2 * 1.74, rel=1e-6)
        else:
            assert list(res.normal) == [0.0, 0.0, 0.0]


def _chain(path, leaf_value=7, normal=955):
    """Interior chain following `path` (child index per level); the last level holds one solid surface leaf."""
    target = S.surface(leaf_value, normal)
    node = None
    for depth in reversed(range(len(path))):
        if depth == len(path) - 1:
            ch = [target if i == path[depth] else S.nonsurf(0) for i in range(8)]
        else:
            ch = [node if i == path[depth] else S.subdiv(0) for i in range(8)]
        node = S.interior(1, ch)
    return node, target


def test_kat4_depth13_single_voxel(oracle):
    path = [7, 0, 5, 2, 7, 0, 3, 4, 1, 6, 7, 0, 5]
    root, target = _chain(path)
    nodes = S.serialise(root)
    lo, size = [1.0, 1.0, 1.0], 1.0
    for k in path:
        lo, hi = child_box(k, lo, size)
        size /= 2
    assert size == 2.0 ** -13
    for dvec in ((0.4, -0.3, -1.0), (-1.0, 0.2, 0.5), (0.3, 1.0, -0.2)):
        o, d = aim(lo, hi, dvec, back=0.9)
        hit, res, st = oracle.cast(nodes, o, d, max_depth=13)
        assert hit and res.pointer == target.offset and res.value == 7
        assert res.depth == 13 and res.scale == 2.0 ** -13
        assert res.t == pytest.approx(S.slab_hit(o, d, lo, hi), rel=1e-5)
        assert st.stale_pops == 0
        # MAX_DEPTH 9 stops at the depth-9 interior ancestor (svotrace.comp:300-302); its child-type mask is then
        # decoded as a "normal" (the reference's quirk): value is the interior's own value
        hit9, res9, _ = oracle.cast(nodes, o, d, max_depth=9)
        assert hit9 and res9.depth == 9 and res9.value == 1 and res9.scale == 2.0 ** -9


def test_kat5_axis_parallel_ray(oracle):
    nodes = root_with([S.nonsurf(2 if i == 0 else 0) for i in range(8)])
    hit, res, _ = oracle.cast(nodes, (1.25, 1.25, 3.0), (0.0, 0.0, -1.0))
    assert hit and res.pointer == 7 and res.t == 1.5  # through child 4 (empty) into child 0: z from 3.0 to 1.5
    # Upstream quirk, reproduced: with dir.y == 0 the y coefficient is -inf and every y plane test is NaN (false), so
    # the traversal never leaves the low-y half whatever origin.y is -- this ray "hits" child 0 although it
    # geometrically passes through child 2's column (svotrace.comp:226-257: EPSILON * sign(0) = 0).
    hit, res, _ = oracle.cast(nodes, (1.25, 1.75, 3.0), (0.0, 0.0, -1.0))
    assert hit and res.pointer == 7 and res.t == 1.5


def test_kat6_nan_direction_hits_the_cap(oracle):
    nodes = root_with([S.nonsurf(2) for _ in range(8)])
    nan = float("nan")
    hit, res, st = oracle.cast(nodes, (1.5, 1.5, 3.0), (nan, nan, nan))
    assert not hit and st.capped == 1 and st.iters == 1500
    assert list(res.debugColor) == pytest.approx([0.3, 0.3, 0.6])  # set on entry (svotrace.comp:213), kept by the cap exit


def test_kat7_packed_normals(oracle):
    for packed, want in ((955, (0, 0, 1)), (551, (-1, 0, 0)), (595, (0, 1, 0)), (999, (1 / math.sqrt(3),) * 3)):
        nodes = root_with([S.surface(1, packed) for _ in range(8)])
        hit, res, _ = oracle.cast(nodes, (1.3, 1.2, 3.0), (0.05, 0.1, -1.0))
        assert hit and list(res.normal) == pytest.approx(list(want), abs=1e-6)
    nodes = root_with([S.surface(1, 555) for _ in range(8)])
    hit, res, _ = oracle.cast(nodes, (1.3, 1.2, 3.0), (0.05, 0.1, -1.0))
    assert hit and all(math.isnan(v) for v in res.normal)  # normalize(vec3(0)) = 0/0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_kat8_random_worlds_vs_bruteforce_geometry(oracle, seed):
    rng = np.random.default_rng(seed)
    n = 32
    vox = np.zeros((n, n, n), np.uint8)
    idx = rng.integers(0, n, size=(300, 3))
    vox[idx[:, 2], idx[:, 1], idx[:, 0]] = rng.integers(1, 4, size=300)
    vox[4:12, 4:12, 4:12] = 2  # a big homogeneous block (subdividable leaves inside)
    nodes, _ = oracle.build_dense(vox)
    assert np.array_equal(S.decode_voxels(nodes, n), vox)
    zs, ys, xs = np.nonzero(vox)
    lo = 1.0 + np.stack([xs, ys, zs], 1) / n
    hi = lo + 1.0 / n
    nr = 1500
    rays = np.zeros(nr, dtype=oracle.RAY_DTYPE)
    o = rng.uniform(0.6, 2.4, (nr, 3))
    tgt = rng.uniform(1.0, 2.0, (nr, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["o"], rays["d"] = o.astype(np.float32), d.astype(np.float32)
    got, st = oracle.cast_rays(nodes, rays, max_depth=5)
    assert st.stale_pops == 0
    o64, d64 = rays["o"].astype(np.float64), rays["d"].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (lo[None] - o64[:, None]) / d64[:, None]
        t1 = (hi[None] - o64[:, None]) / d64[:, None]
    tn = np.minimum(t0, t1).max(axis=2)
    tf = np.maximum(t0, t1).min(axis=2)
    tn = np.maximum(tn, 0.0)
    ok = tn <= tf
    # ignore grazing contacts (chord shorter than 1e-4): float32 traversal may legitimately differ there
    solid_chord = np.where(ok, tf - tn, 0.0)
    tfirst = np.where(ok & (solid_chord > 1e-4), tn, np.inf).min(axis=1)
    tany = np.where(ok, tn, np.inf).min(axis=1)
    hit = got["id"] != oracle.NO_HIT
    clear_hit = np.isfinite(tfirst) & (tfirst == tany)
    clear_miss = ~np.isfinite(tany)
    assert hit[clear_hit].all() and not hit[clear_miss].any()
    rel = np.abs(got["t"][clear_hit] - tfirst[clear_hit]) / np.maximum(tfirst[clear_hit], 1e-3)
    assert rel.max() < 1e-4
    assert clear_hit.sum() > 200 and clear_miss.sum() > 50


def test_oracle_math_accuracy(oracle):
    """The fixed polynomial kernels are within a few ulp of the true functions on the ranges the shader uses."""
    L = oracle.lib()
    rng = np.random.default_rng(3)

    def ulps(got, want):
        want32 = want.astype(np.float32)
        return np.abs(got.astype(np.float64) - want) / np.maximum(np.spacing(np.abs(want32)).astype(np.float64), 1e-45)

    x = rng.uniform(-7, 7, 4000).astype(np.float32)
    s = np.array([L.svo_oracle_sin(float(v)) for v in x], np.float32)
    c = np.array([L.svo_oracle_cos(float(v)) for v in x], np.float32)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 2e-7 and np.abs(c - np.cos(x.astype(np.float64))).max() < 2e-7
    big = rng.uniform(0, 8e5, 4000).astype(np.float32)  # rand()'s argument range
    s = np.array([L.svo_oracle_sin(float(v)) for v in big], np.float32)
    assert np.abs(s - np.sin(big.astype(np.float64))).max() < 2e-7
    a = rng.uniform(-1, 1, 4000).astype(np.float32)
    ac = np.array([L.svo_oracle_acos(float(v)) for v in a], np.float32)
    assert np.abs(ac - np.arccos(a.astype(np.float64))).max() < 1e-6
    e = rng.uniform(-20, 3, 4000).astype(np.float32)
    ex = np.array([L.svo_oracle_exp(float(v)) for v in e], np.float32)
    assert ulps(ex, np.exp(e.astype(np.float64))).max() < 4
    r = np.array([L.svo_oracle_rand(float(p), float(q)) for p, q in zip(big[:500], big[500:1000])], np.float32)
    assert ((r >= 0) & (r < 1)).all()


def test_iteration_cap_boundary(oracle):
    """svotrace.comp:263-266: `iter++; if (iter > 1500) return false`.  A ray that leaves the cube during iteration
    1500 is an ordinary miss with iter = 1500; one that would need iteration 1501 is capped (iter = 1501, the entry
    colour kept).  Long walks through an all-empty subdivided row give both kinds."""
    import svo_stream as S
    nodes = S.serialise(S.tube(10))
    rays = S.tube_rays(10, 20000)
    out, st = oracle.cast_rays(nodes, rays, 13, nthreads=4)
    assert (out["id"] == 0xFFFFFFFF).all()
    it = out["iter"]
    assert it.max() == 1501 and st.capped == (it == 1501).sum() > 0
    assert (it == 1500).sum() > 0 and ((it > 1480) & (it < 1500)).sum() > 0
    # one of each through the single-cast entry point: debugColor tells the two exits apart
    i_cap, i_last = int(np.flatnonzero(it == 1501)[0]), int(np.flatnonzero(it == 1500)[0])
    hit, res, s1 = oracle.cast(nodes, rays["o"][i_cap], rays["d"][i_cap])
    assert not hit and s1.capped == 1 and list(res.debugColor) == pytest.approx([0.3, 0.3, 0.6])
    hit, res, s1 = oracle.cast(nodes, rays["o"][i_last], rays["d"][i_last])
    assert not hit and s1.capped == 0 and s1.iters == 1500 and list(res.debugColor) == pytest.approx([15.0, 15.0, 15.0])
