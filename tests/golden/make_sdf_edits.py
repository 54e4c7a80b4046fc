#!/usr/bin/env python3
"""Makes tests/golden/sdf_edits.npz: a brush session of the REFERENCE'S OWN edit code on the 128^3 test terrain.

Octree.useSDFBrush / subdivideNode / the ChangeBounds bookkeeping (Octree.java:676-885) and sdf/Sphere.java, sdf/Box.java are
compiled for the CPU from /root/reference by oracle/build_ref_java.py; this script plays Main.placeSDF (Main.java:338-353)
with them -- additive and subtractive spheres and a box, each applied to the result of the one before -- and records what the
engine would push with Renderer.updateSSBO: per step the two byte ranges and the bytes inside them, plus the new memOffset.
/root/reference does not exist on the GPU box: the -m gpu test replays these ranges through svo_upload_range.

Usage: python tests/golden/make_sdf_edits.py      (needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import ref_java as RJ  # noqa: E402
import svo_raytracer_b200 as svo  # noqa: E402

N, CHUNK, MAX_LOD = 128, 64, 7


def session():
    """(kind, origin, params, value) of every stroke; positions relative to the terrain's surface so that they touch it."""
    hm, mm = svo.terrain_inputs(N)
    hs = (hm.astype(np.uint32) * (N // 4)) >> 16
    s = lambda x, z: int(hs[z, x])  # noqa: E731
    return [("sphere", (40, s(40, 50) + 3, 50), (6,), 1),       # a mound on the surface
            ("sphere", (40, s(40, 50) - 2, 50), (5,), 0),       # dig into it
            ("sphere", (90, s(90, 30) + 10, 30), (4,), 2),      # a ball in the air (new nodes only)
            ("box", (70, s(70, 80) + 2, 80), (5, 3, 4), 3),     # a slab
            ("sphere", (64, 100, 64), (3,), 0),                 # subtract where nothing is: no change at all
            ("sphere", (41, s(40, 50) + 1, 52), (7,), 0),       # a crater over the first two strokes
            ("sphere", (20, s(20, 20), 20), (9,), 1)]           # a big additive stroke half inside the ground


def main():
    hm, mm = svo.terrain_inputs(N)
    nodes, _ = O.build_terrain(hm, mm, N, CHUNK)
    out = {"n": np.int32(N), "chunk": np.int32(CHUNK), "max_lod": np.int32(MAX_LOD), "steps": np.int32(len(session())),
           "base_bytes": np.int64(nodes.size)}
    cur = nodes
    for k, (kind, origin, params, value) in enumerate(session()):
        new, ranges, cb = RJ.sdf_brush(cur, N, MAX_LOD, origin, params, value, kind=kind)
        # what the engine pushes with Renderer.updateSSBO: the bytes of [start0, end0) and [start1, end1).  Stored sparsely: the
        # bytes below the old memOffset that changed inside the ranges (position, value) and the appended tail; the test rebuilds
        # the ranges' contents from the stream it holds.
        out["s%d_bounds" % k] = np.asarray(cb, np.int64)
        out["s%d_what" % k] = np.asarray([0 if kind == "sphere" else 1, value, *origin, *params], np.int32)
        inside = np.zeros(new.size, bool)
        for a, b in ranges:
            inside[a:b] = True
        m = cur.size
        changed = np.nonzero(cur != new[:m])[0]
        pushed = changed[inside[changed]]
        stale = changed[~inside[changed]]  # the engine's own copy changes here too, its GPU copy does not (markNodeAsDirty, Octree.java:787)
        assert all(new[i] == 127 for i in stale), "bytes changed outside the ChangeBounds that are not DELETE_VALUE marks"
        assert new.size == m or (len(ranges) > 0 and ranges[-1] == (m, new.size)), "the appended records are the second range"
        out["s%d_idx" % k] = pushed.astype(np.int32)
        out["s%d_val" % k] = new[pushed].copy()
        out["s%d_tail" % k] = new[m:].copy()
        out["s%d_stale" % k] = stale.astype(np.int32)
        print("step %d: %s value %d at %s: %d -> %d bytes, ranges %s, %d stale bytes" % (k, kind, value, origin, cur.size, new.size, ranges, stale.size))
        cur = new
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sdf_edits.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
