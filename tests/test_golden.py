"""Committed golden vectors (tests/golden/svo_golden.npz, made by tests/golden/make_golden.py).

The frames, hit records and beam images in the file are OUTPUTS OF THE REFERENCE'S OWN SHADERS (svotrace.comp /
svobeam.comp compiled for the CPU: oracle/build_ref.py -> oracle/_ref/libsvo_ref.so), generated in the build
container where /root/reference exists.  CPU tests: the restated oracle, the product's world generator and its
terrain generator reproduce them.  GPU tests: the CUDA path, through the C ABI, reproduces them with neither the
oracle nor /root/reference in the loop -- this is how the GPU box checks against reference-derived outputs."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "svo_golden.npz"))
N, CHUNK, W, H, DEPTH = (int(v) for v in G["params"])
KEYS = [c + str(m) for c in "ABC" for m in (0, 1, 2, 3, 4)] + [c + str(m) + "s" for c in "ABC" for m in (0, 2)]


def _depth_of(key):
    return 13 if key.endswith("s") else DEPTH  # "s": as shipped, MAX_DEPTH 13 (svotrace.comp:40)


def test_fixtures_come_from_the_reference_shaders():
    assert "svotrace.comp" in str(G["source"]) and "svobeam.comp" in str(G["source"])


def same_f32_bits(a_bits, b_bits):
    """Bit equality of float planes stored as uint32, with every NaN equal to every NaN (the NaN payload and sign that
    0/0 produces differ between x86 SSE and the GPU; GLSL gives NaNs no payload semantics either)."""
    a, b = np.asarray(a_bits, np.uint32), np.asarray(b_bits, np.uint32)
    an, bn = np.isnan(a.view(np.float32)), np.isnan(b.view(np.float32))
    return bool(np.array_equal(an, bn) and np.array_equal(a[~an], b[~bn]))


def test_inputs_and_builders_reproduce_golden(svo, oracle):
    hm, mm = svo.terrain_inputs(N, seed=1)
    assert np.array_equal(hm, G["height"]) and np.array_equal(mm, G["mat"])
    nodes, counts = oracle.build_terrain(hm, mm, N, CHUNK)
    assert np.array_equal(nodes, G["nodes"]) and list(counts) == list(G["counts"])
    assert np.array_equal(svo.build_terrain(hm, mm, N, CHUNK), G["nodes"])


@pytest.mark.parametrize("key", KEYS)
def test_oracle_reproduces_golden_frames(svo, oracle, key):
    cam, mode = key[0], int(key[1])
    pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
    planes, st = oracle.render(G["nodes"], oracle.make_frame(pos, l1, l2, r1, r2, frame_number=5, render_mode=mode, max_depth=_depth_of(key)), W, H)
    assert np.array_equal(planes["rgba8"], G[key + "_rgba8"])
    assert same_f32_bits(planes["depth"].view(np.uint32), G[key + "_depth"])
    assert same_f32_bits(planes["radiance"].view(np.uint32), G[key + "_radiance"])
    assert np.array_equal(planes["hit_id"], G[key + "_hit_id"]) and np.array_equal(planes["iter"], G[key + "_iter"])
    assert [st.casts, st.iters, st.record_bytes] == list(G[key + "_stats"])


def test_oracle_reproduces_golden_ray_stream(oracle):
    rays = np.ascontiguousarray(G["rays"]).view(oracle.RAY_DTYPE).reshape(-1)
    hits, _ = oracle.cast_rays(G["nodes"], rays, max_depth=DEPTH)
    assert np.array_equal(hits["id"], G["hits_id"]) and np.array_equal(hits["iter"], G["hits_iter"])
    assert np.array_equal(hits["value"], G["hits_value"]) and same_f32_bits(hits["t"].view(np.uint32), G["hits_t"])


@pytest.mark.parametrize("cam", ["A", "C"])
def test_oracle_reproduces_golden_beam(svo, oracle, cam):
    pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
    f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=5, render_mode=2, use_beam=1, max_depth=5)
    b = oracle.beam(G["beam_nodes"], f, W, H)
    assert same_f32_bits(b.view(np.uint32), G["beam_%s" % cam])
    fine, _ = oracle.render(G["beam_nodes"], f, W, H, beam=b)
    assert np.array_equal(fine["rgba8"], G["beam_%s_rgba8" % cam])
    assert same_f32_bits(fine["depth"].view(np.uint32), G["beam_%s_depth" % cam])


@pytest.mark.gpu
@pytest.mark.parametrize("cam", ["A", "C"])
def test_cuda_reproduces_golden_beam(svo, cam):
    with svo.SvoContext(W, H) as c:
        c.upload(G["beam_nodes"])
        f = svo.camera_frame(cam, frame_number=5, render_mode=2, use_beam=1, max_depth=5)
        c.beam(f)
        assert same_f32_bits(c.read_plane(svo._lib.PLANE_BEAM).view(np.uint32), G["beam_%s" % cam])
        c.render(f)
        assert np.array_equal(c.read_color_rgba8(), G["beam_%s_rgba8" % cam])
        assert same_f32_bits(c.read_depth().view(np.uint32), G["beam_%s_depth" % cam])


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [0, 1, 2, 6, 10])
def test_cuda_reproduces_golden(svo, kernel):
    with svo.SvoContext(W, H) as c:
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.upload(G["nodes"])
        for key in KEYS:
            c.render(svo.camera_frame(key[0], frame_number=5, render_mode=int(key[1]), max_depth=_depth_of(key)))
            assert np.array_equal(c.read_color_rgba8(), G[key + "_rgba8"]), key
            assert same_f32_bits(c.read_depth().view(np.uint32), G[key + "_depth"]), key
            assert same_f32_bits(c.read_radiance().view(np.uint32), G[key + "_radiance"]), key
            assert np.array_equal(c.read_hit_id(), G[key + "_hit_id"]) and np.array_equal(c.read_iter(), G[key + "_iter"]), key
            st = c.render_stats(svo.camera_frame(key[0], frame_number=5, render_mode=int(key[1]), max_depth=_depth_of(key)))
            assert [st["casts"], st["iters"], st["record_bytes"]] == list(G[key + "_stats"]), key
        rays = np.ascontiguousarray(G["rays"]).view(svo.RAY_DTYPE).reshape(-1)
        hits = c.cast(rays, max_depth=DEPTH)
        assert np.array_equal(hits["id"], G["hits_id"]) and np.array_equal(hits["iter"], G["hits_iter"])
        assert np.array_equal(hits["value"], G["hits_value"]) and same_f32_bits(hits["t"].view(np.uint32), G["hits_t"])
