"""Incremental transcode (svo_upload_range -> gpu_diff_apply + gpu_patch, csrc/svo_gpu_transcode.cu) on the SIMT emulator:
after edits made the way the engine's SDF brush makes them (Octree.java:700-885) and pushed as the byte ranges
Renderer.updateSSBO would push (Main.java:349-350), the patched descriptor tree equals a whole transcode of the edited
stream (layout-independent fingerprint), the content boxes cover it, and the work done is proportional to the edit.
The same scenarios run on the B200 through the C ABI in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import svo_stream as S
from hostemu import emu as E


def _check(old, ed, ranges=None, expect_whole=False):
    new = ed.stream()
    r = E.patch_check(old, new, ed.ranges() if ranges is None else ranges)
    if expect_whole:
        assert r["fell_back"] >= 1, r
    else:
        assert r["status"] == 0 and r["fell_back"] == 0, r
    return r


def test_whole_transcode_on_emulator_equals_host(terrain128):
    assert E.gpu_transcode_check(terrain128) == 0
    import svo_stream
    assert E.gpu_transcode_check(svo_stream.serialise(svo_stream.tube(6))) == 0
    assert E.gpu_transcode_check(np.zeros(0, np.uint8)) == 0
    assert E.gpu_transcode_check(np.zeros(7, np.uint8)) == 0


def test_single_value_edits(terrain128):
    rng = np.random.default_rng(1)
    for code, nonzero in ((1, True), (3, True), (2, False), (2, True)):
        ed = S.StreamEditor(terrain128)
        try:
            path = ed.find_leaf(code, nonzero, rng)
        except AssertionError:
            continue
        ed.set_value(path, 0 if nonzero else 3)
        r = _check(terrain128, ed)
        assert 1 <= r["dirty"] <= 2 and r["roots"] == 1, r  # the parent's descriptor (value != 0 mask)
        assert r["appended"] <= 64, r


def test_noop_range_costs_nothing(terrain128):
    ed = S.StreamEditor(terrain128)
    r = E.patch_check(terrain128, terrain128, [(1000, 9000), (terrain128.size - 500, terrain128.size)])
    assert r == {"status": 0, "dirty": 0, "roots": 0, "appended": 0, "fell_back": 0, "reachable": r["reachable"], "stored": r["stored"]}
    assert r["stored"] == r["reachable"]


def test_subdivide_in_the_air_grows_the_content_box(terrain128):
    """New voxels where nothing was: an empty subdividable leaf becomes interior, its children are appended (two levels)."""
    rng = np.random.default_rng(2)
    ed = S.StreamEditor(terrain128)
    path = ed.find_leaf(2, False, rng, min_depth=3)
    ed.subdivide(path, [0, 2, 0, 0, 3, 0, 0, 1])
    ed.subdivide(path + [1], [1, 1, 0, 0, 0, 0, 2, 2])
    ed.subdivide(path + [1, 7], [3] * 8, surface=True)
    r = _check(terrain128, ed)
    assert r["roots"] >= 1 and 1 <= r["appended"] <= 200, r
    assert len(ed.ranges()) == 2


def test_ranges_in_either_order_and_one_by_one(terrain128):
    rng = np.random.default_rng(3)
    ed = S.StreamEditor(terrain128)
    path = ed.find_leaf(2, True, rng, min_depth=2)   # a buried solid block
    ed.subdivide(path, [1, 0, 1, 0, 2, 0, 2, 0])       # carve half of it away
    r0, r1 = ed.ranges()
    _check(terrain128, ed, [r0, r1])
    _check(terrain128, ed, [r1, r0])
    # the first push alone leaves a tree whose new children read as zeros: still a valid tree, completed by the second push
    half = E.patch_check(terrain128, ed.stream()[:r1[0]], [r0]) if r0[1] <= r1[0] else None
    assert half is None or half["status"] == 0


def test_collapse_and_many_scattered_edits(terrain128):
    rng = np.random.default_rng(4)
    ed = S.StreamEditor(terrain128)
    for _ in range(25):
        kind = int(rng.integers(0, 3))
        try:
            if kind == 0:
                ed.set_value(ed.find_leaf(1, True, rng), int(rng.integers(0, 4)))
            elif kind == 1:
                p = ed.find_leaf(2, bool(rng.integers(0, 2)), rng, min_depth=3)
                ed.subdivide(p, [int(v) for v in rng.integers(0, 4, 8)])
            else:
                # collapse a deep interior node
                off, path, code = 0, [], 0
                for _d in range(int(rng.integers(3, 6))):
                    n = int(rng.integers(0, 8))
                    o2, c2 = ed.child(off, n)
                    if c2 != 0 or ed.cp(o2) == 0:
                        break
                    off, code = o2, c2
                    path.append(n)
                if len(path) >= 3:
                    ed.collapse(path, int(rng.integers(0, 3)))
        except AssertionError:
            pass
    r = _check(terrain128, ed)
    assert r["roots"] >= 5 and r["stored"] > r["reachable"] - 1, r


def test_edit_at_the_top_of_the_tree_falls_back_to_a_whole_transcode(terrain128):
    ed = S.StreamEditor(terrain128)
    ed.set_value([3], 0 if ed.buf[ed.child(0, 3)[0]] else 1)  # a child of the root: the root is the re-walk's root
    _check(terrain128, ed, expect_whole=True)


def test_hand_built_tree_every_record_type():
    kids = [S.surface(1, normal=0), S.nonsurf(2), S.subdiv(3), S.surface(2, normal=555), S.nonsurf(0), S.surface(3, normal=999),
            S.interior(1, [S.surface(1, normal=123), S.nonsurf(0), S.nonsurf(1), S.subdiv(0), S.surface(2, normal=987), S.nonsurf(0),
                           S.nonsurf(0), S.surface(3, normal=505)]), S.subdiv(0)]
    nodes = S.serialise(S.interior(1, kids))
    ed = S.StreamEditor(nodes)
    ed.subdivide([6, 3], [1, 2, 3, 0, 0, 0, 0, 1])
    ed.set_value([6, 0], 0)
    r = E.patch_check(nodes, ed.stream(), ed.ranges())
    assert r["status"] == 0, r
