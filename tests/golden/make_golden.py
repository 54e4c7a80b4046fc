#!/usr/bin/env python
"""Regenerates tests/golden/svo_golden.npz.

PARITY UNPINNED BY THE REFERENCE: dyoo47/svo-raytracer ships no test vector for its traversal and neither its
Java host nor its GLSL can run in this image, so these vectors are produced by the CPU oracle (oracle/svo_oracle.c,
itself pinned by tests/test_oracle_kat.py).  They freeze today's oracle + builder + terrain generator so that any
later drift of either side is caught, and they let the GPU box check the CUDA path without /root/reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import svo_raytracer_b200 as svo  # noqa: E402
from oracle import oracle as O  # noqa: E402

N, CHUNK, W, H, DEPTH = 64, 32, 96, 54, 6


def main():
    hm, mm = svo.terrain_inputs(N, seed=1)
    nodes, counts = O.build_terrain(hm, mm, N, CHUNK)
    out = {"nodes": nodes, "height": hm, "mat": mm, "counts": np.array(counts, np.uint64),
           "params": np.array([N, CHUNK, W, H, DEPTH], np.int32)}
    for cam in "ABC":
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        for mode in (0, 2, 3):
            f = O.make_frame(pos, l1, l2, r1, r2, frame_number=5, render_mode=mode, max_depth=DEPTH)
            planes, st = O.render(nodes, f, W, H)
            key = "%s%d_" % (cam, mode)
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
    hits, _ = O.cast_rays(nodes, rays, max_depth=DEPTH)
    out["rays"] = rays.view(np.float32).reshape(n, 6)
    out["hits_id"], out["hits_iter"], out["hits_value"] = hits["id"], hits["iter"], hits["value"]
    out["hits_t"] = hits["t"].view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "svo_golden.npz"), **out)
    print("wrote svo_golden.npz:", {k: v.shape for k, v in out.items() if k in ("nodes", "rays")})


if __name__ == "__main__":
    main()
