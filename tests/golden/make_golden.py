#!/usr/bin/env python
"""Regenerates tests/golden/svo_golden.npz FROM THE REFERENCE'S OWN SHADERS.

The frames, hit records and beam image in the file are outputs of oracle/_ref/libsvo_ref.so -- svotrace.comp and
svobeam.comp from /root/reference compiled for the CPU by oracle/build_ref.py (g++ through oracle/glsl_shim.h) --
NOT of the restatement oracle/svo_oracle.c.  /root/reference does not exist on the GPU box, so these committed
vectors are how the CUDA path is checked against reference-derived outputs there (tests/test_golden.py).

Inputs (node stream) come from the restated builder, oracle/svo_builder.c: Octree.java cannot be compiled here
(no JVM).  The `*_stats` entries (casts / iterations / record bytes) are the oracle's instrumentation -- the shader
has no such counters; they are cross-checked against the reference's iteration plane in main() below.

    python tests/golden/make_golden.py        (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import svo_raytracer_b200 as svo  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402

N, CHUNK, W, H, DEPTH = 64, 32, 96, 54, 6
MODES = (0, 1, 2, 3, 4)


def beam_world(n=32, seed=3):
    """No subdividable leaves (svobeam.comp's extractChild has no case for them): see tests/test_oracle_ref.py."""
    rng = np.random.default_rng(seed)
    vox = (rng.random((n, n, n)) < 0.12).astype(np.uint8) * rng.integers(1, 4, (n, n, n)).astype(np.uint8)
    z, y, x = np.mgrid[0:n, 0:n, 0:n]
    vox[y < n // 3] = 1
    vox[(x % 2 == 0) & (y % 2 == 0) & (z % 2 == 0)] = 0
    vox[(x % 2 == 1) & (y % 2 == 0) & (z % 2 == 0)] = 2
    nodes, counts = O.build_dense(vox)
    assert counts[2] == 0
    return nodes


def main():
    assert os.path.isdir("/root/reference"), "golden vectors are generated from the reference's shaders"
    R.lib()
    hm, mm = svo.terrain_inputs(N, seed=1)
    nodes, counts = O.build_terrain(hm, mm, N, CHUNK)
    out = {"nodes": nodes, "height": hm, "mat": mm, "counts": np.array(counts, np.uint64),
           "params": np.array([N, CHUNK, W, H, DEPTH], np.int32),
           "source": np.array(R.lib().svo_ref_about().decode())}
    for cam in "ABC":
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        for mode in MODES:
            for tag, depth in (("", DEPTH), ("s", 13)):  # "s" = as shipped: MAX_DEPTH 13 on the 6-level tree
                if tag and mode not in (0, 2):
                    continue
                f = O.make_frame(pos, l1, l2, r1, r2, frame_number=5, render_mode=mode, max_depth=depth)
                planes = R.render(nodes, f, W, H)
                chk, st = O.render(nodes, f, W, H)
                assert np.array_equal(chk["iter"], planes["iter"])  # the stats below describe the same casts
                key = "%s%d%s_" % (cam, mode, tag)
                out[key + "rgba8"] = planes["rgba8"]
                out[key + "depth"] = planes["depth"].view(np.uint32)
                out[key + "radiance"] = planes["radiance"].view(np.uint32)
                out[key + "hit_id"] = planes["hit_id"]
                out[key + "iter"] = planes["iter"]
                out[key + "stats"] = np.array([st.casts, st.iters, st.record_bytes], np.uint64)
    rng = np.random.default_rng(2026)
    n = 3000
    rays = np.zeros(n, dtype=O.RAY_DTYPE)
    rays["o"] = rng.uniform(0.9, 2.1, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["d"][:8, 1] = 0.0
    rays["d"][8:12] = np.nan
    hits = R.cast_rays(nodes, rays, max_depth=DEPTH)
    out["rays"] = rays.view(np.float32).reshape(n, 6)
    out["hits_id"], out["hits_iter"], out["hits_value"] = hits["id"], hits["iter"], hits["value"]
    out["hits_t"] = hits["t"].view(np.uint32)
    # beam pre-pass (svobeam.comp) and the fine pass that consumes it
    bn = beam_world()
    out["beam_nodes"] = bn
    for cam in "AC":
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        f = O.make_frame(pos, l1, l2, r1, r2, frame_number=5, render_mode=2, use_beam=1, max_depth=5)
        b = R.beam(bn, f, W, H)
        fine = R.render(bn, f, W, H, beam=b)
        out["beam_%s" % cam] = b.view(np.uint32)
        out["beam_%s_rgba8" % cam] = fine["rgba8"]
        out["beam_%s_depth" % cam] = fine["depth"].view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "svo_golden.npz"), **out)
    print("wrote svo_golden.npz:", {k: v.shape for k, v in out.items() if k in ("nodes", "rays", "beam_nodes")}, len(out), "arrays")


if __name__ == "__main__":
    main()
