"""World generation on the device (svo_build_terrain_device, csrc/svo_gpu_build.cu) on the B200: the node stream built
by kernels straight into HBM equals the host builder's byte for byte (which equals the brute-force restatement of
Octree.java, tests/test_builder.py), the scene it leaves in the context renders like an uploaded one, and the
8192^3 bench world builds in well under a second."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,chunk", [(8, 4), (64, 16), (128, 128), (256, 64), (512, 512), (1024, 1024), (2048, 1024)])
def test_device_builder_equals_host_builder(svo, n, chunk):
    for seed in ((1, 7) if n <= 512 else (1,)):
        hm, mm = svo.terrain_inputs(n, seed=seed)
        want = svo.build_terrain(hm, mm, n, chunk)
        with svo.SvoContext(64, 64) as c:
            nbytes = c.build_terrain_device(hm, mm, n, chunk)
            assert nbytes == want.size
            got = c.download()
            assert np.array_equal(got, want), (n, chunk, seed, int(np.flatnonzero(got != want)[0]))
            probe_built = c.scene_probe()
            c.upload(want)
            assert c.scene_probe() == probe_built  # same descriptors, reference offsets and content boxes


def test_device_builder_reproduces_the_reference_builders_streams(svo):
    """tests/golden/builder_golden.json: SHA-256 of the streams the reference's OWN builder produces (Octree.java, OctreeThread.java
    and the voxeliser shader compiled from their text, tests/golden/make_builder_golden.py) -- 8^3 ... 1024^3, the last with
    the constants as shipped.  The device builder's bytes have those digests."""
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "builder_golden.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 8
    with svo.SvoContext(64, 64) as c:
        for case in cases:
            hm, mm = svo.terrain_inputs(case["n"], seed=case["seed"])
            assert c.build_terrain_device(hm, mm, case["n"], case["chunk"]) == case["bytes"], case
            assert hashlib.sha256(c.download().tobytes()).hexdigest() == case["sha256"], case


def test_device_builder_adversarial_maps(svo):
    from test_builder import _adversarial_maps
    n, chunk = 64, 32
    with svo.SvoContext(64, 64) as c:
        for hm, mm in _adversarial_maps(n, np.random.default_rng(11)):
            c.build_terrain_device(hm, mm, n, chunk)
            assert np.array_equal(c.download(), svo.build_terrain(hm, mm, n, chunk))


def test_device_built_scene_renders_like_an_uploaded_one(svo, oracle):
    n, chunk, W, H = 256, 128, 320, 180
    hm, mm = svo.terrain_inputs(n)
    with svo.SvoContext(W, H) as c:
        c.build_terrain_device(hm, mm, n, chunk)
        nodes = c.download()
        for cam, mode in (("B", 0), ("C", 2)):
            pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
            want, _ = oracle.render(nodes, oracle.make_frame(pos, l1, l2, r1, r2, frame_number=2, render_mode=mode, max_depth=8), W, H,
                                    nthreads=8, planes=("rgba8", "depth"))
            c.render(svo.camera_frame(cam, frame_number=2, render_mode=mode, max_depth=8))
            assert np.array_equal(c.read_color_rgba8(), want["rgba8"])
            assert np.array_equal(c.read_depth().view(np.uint32), want["depth"].view(np.uint32))


def test_device_builder_refuses_what_it_does_not_take(svo):
    hm, mm = svo.terrain_inputs(64)
    with svo.SvoContext(64, 64) as c:
        with pytest.raises(svo.SvoError) as e:
            c.build_terrain_device(hm, mm, 64, 2)  # chunk of 2 voxels: sub-octrees without levels
        assert e.value.code == svo._lib.ERR_INVALID


def test_bench_world_builds_on_the_device_in_under_a_second(svo):
    """8192^3 (BASELINE's headline world, 1.95 GB of nodes): byte-equal to the host builder, and timed."""
    n, chunk = 8192, 1024
    hm, mm = svo.terrain_inputs(n)
    t0 = time.time()
    want = svo.build_terrain(hm, mm, n, chunk)
    host_s = time.time() - t0
    with svo.SvoContext(64, 64) as c:
        t0 = time.time()
        nbytes = c.build_terrain_device(hm, mm, n, chunk)  # includes the H2D copy of the maps and the descriptor transcode
        dev_s = time.time() - t0
        assert nbytes == want.size
        got = c.download()
        same = np.array_equal(got, want)
        print("8192^3 world: host builder %.2f s, device builder + transcode %.3f s, %d bytes" % (host_s, dev_s, nbytes))
        assert same
        assert dev_s < 1.5
