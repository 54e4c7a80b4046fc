#!/usr/bin/env python3
"""Makes tests/golden/builder_golden.json: SHA-256 of the node streams that THE REFERENCE'S OWN BUILDER (Octree.java,
OctreeThread.java and chunkgen-heightmap.comp compiled from their text by oracle/build_ref_java.py) produces for the test
terrains.  /root/reference does not exist on the GPU box: there tests/test_gpu_build.py holds the device builder to these
digests directly.  The 1024^3 case runs the reference's constants as shipped (CHUNK_SIZE 1024, 512^3 OctreeThreads, maxLOD 9).

Usage: python tests/golden/make_builder_golden.py      (needs /root/reference; ~15 s)
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_java as RJ  # noqa: E402
import svo_raytracer_b200 as svo  # noqa: E402

CASES = [(8, 4, 1), (64, 16, 1), (64, 16, 7), (128, 64, 1), (128, 128, 7), (256, 64, 1), (512, 512, 1), (1024, 1024, 1)]


def main():
    out = {"what": "sha256 of the reference builder's node stream for svo.terrain_inputs(n, seed), chunk", "cases": []}
    for n, chunk, seed in CASES:
        hm, mm = svo.terrain_inputs(n, seed=seed)
        nodes, counts = RJ.build_terrain(hm, mm, n, chunk, cap=max(1 << 20, 64 * n * n))
        out["cases"].append({"n": n, "chunk": chunk, "seed": seed, "bytes": int(nodes.size), "sha256": hashlib.sha256(nodes.tobytes()).hexdigest(),
                             "counts": counts})
        print(n, chunk, seed, nodes.size, out["cases"][-1]["sha256"][:16])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "builder_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path)


if __name__ == "__main__":
    main()
