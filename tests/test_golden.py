"""Committed golden vectors (tests/golden/svo_golden.npz, made by tests/golden/make_golden.py).

Parity is UNPINNED by the reference (it has no fixture for this path); the vectors come from the CPU oracle and
freeze it.  CPU tests: the oracle, the product's world generator and its terrain generator still reproduce them.
GPU test: the CUDA path, through the C ABI, reproduces them without the oracle in the loop."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "svo_golden.npz"))
N, CHUNK, W, H, DEPTH = (int(v) for v in G["params"])
KEYS = [c + str(m) for c in "ABC" for m in (0, 2, 3)]


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
    planes, st = oracle.render(G["nodes"], oracle.make_frame(pos, l1, l2, r1, r2, frame_number=5, render_mode=mode, max_depth=DEPTH), W, H)
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


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [0, 1, 2, 6])
def test_cuda_reproduces_golden(svo, kernel):
    with svo.SvoContext(W, H) as c:
        c.set_option(svo._lib.OPT_AUX_PLANES, 1)
        c.set_option(svo._lib.OPT_KERNEL, kernel)
        c.upload(G["nodes"])
        for key in KEYS:
            c.render(svo.camera_frame(key[0], frame_number=5, render_mode=int(key[1]), max_depth=DEPTH))
            assert np.array_equal(c.read_color_rgba8(), G[key + "_rgba8"]), key
            assert same_f32_bits(c.read_depth().view(np.uint32), G[key + "_depth"]), key
            assert same_f32_bits(c.read_radiance().view(np.uint32), G[key + "_radiance"]), key
            assert np.array_equal(c.read_hit_id(), G[key + "_hit_id"]) and np.array_equal(c.read_iter(), G[key + "_iter"]), key
            st = c.render_stats(svo.camera_frame(key[0], frame_number=5, render_mode=int(key[1]), max_depth=DEPTH))
            assert [st["casts"], st["iters"], st["record_bytes"]] == list(G[key + "_stats"]), key
        rays = np.ascontiguousarray(G["rays"]).view(svo.RAY_DTYPE).reshape(-1)
        hits = c.cast(rays, max_depth=DEPTH)
        assert np.array_equal(hits["id"], G["hits_id"]) and np.array_equal(hits["iter"], G["hits_iter"])
        assert np.array_equal(hits["value"], G["hits_value"]) and same_f32_bits(hits["t"].view(np.uint32), G["hits_t"])
