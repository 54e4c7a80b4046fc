"""Upload-time transcode (host code, no GPU): reference node stream -> breadth-first 8-byte descriptors."""
import ctypes as C
import time

import numpy as np

import svo_stream as S


def probe(svo, nodes, nthreads=1, want_desc=0):
    nodes = np.ascontiguousarray(nodes, np.uint8)
    out = (C.c_uint64 * 8)()
    desc = np.zeros(3 * want_desc, np.uint32) if want_desc else None
    rc = svo._lib.lib().svo_transcode_probe(nodes.ctypes.data_as(C.c_void_p), nodes.size, nthreads, C.byref(out),
                                            None if desc is None else desc.ctypes.data_as(C.c_void_p), 0 if desc is None else desc.size)
    assert rc == 0
    return [int(v) for v in out], (None if desc is None else desc.reshape(-1, 3))


def test_hand_checked_descriptors(svo):
    # root with children: [interior(with children), surface leaf, empty subdiv, nonsurf solid, empty x4]
    grand = [S.nonsurf(1 if i in (0, 7) else 0) for i in range(8)]
    kids = [S.interior(5, grand), S.surface(2, 955), S.subdiv(0), S.nonsurf(3)] + [S.nonsurf(0)] * 4
    nodes = S.serialise(S.interior(1, kids))
    out, desc = probe(svo, nodes, 1, want_desc=2)
    assert out[0] == 2 and out[1] == 2
    codes_root = 0 | (1 << 2) | (2 << 4) | (3 << 6) | (3 << 8) | (3 << 10) | (3 << 12) | (3 << 14)
    # root: first interior child is descriptor 1; value != 0 for children 0,1,3; only child 0 has a descriptor
    assert list(desc[0]) == [1, codes_root | (0b00001011 << 16) | (0b00000001 << 24), 7]
    # the interior child: its 8 one-byte children start right after the root's block (7+3+7+1+4 = 22 bytes)
    assert list(desc[1]) == [2, 0xFFFF | (0b10000001 << 16), 7 + 22]
    # leaf bounds: children 1 and 3 of the root (x in [0.5,1] of the cube) and grandchildren 0 and 7 (x in [0, 0.5])
    assert out[3] == 0 and out[4] == (0 << 32) | (1 << 24)


def test_thread_count_does_not_change_the_result(svo, terrain128, terrain512):
    for nodes in (terrain128, terrain512):
        ref, _ = probe(svo, nodes, 1)
        for t in (2, 5, 16):
            got, _ = probe(svo, nodes, t)
            assert got == ref
        assert ref[0] > 1000 and ref[3] == 0


def test_empty_and_degenerate_streams(svo):
    out, _ = probe(svo, np.zeros(0, np.uint8))
    assert out[0] == 1 and out[3] == 1  # a root read as zeros: nothing to hit
    out, _ = probe(svo, np.zeros(7, np.uint8))
    assert out[0] == 1 and out[3] == 1
    # a root whose child pointer points at itself (cp = 0 reads the root's own bytes as children)
    out, _ = probe(svo, np.array([1, 0, 0, 0, 0, 0, 0], np.uint8))
    assert out[0] >= 1


def cyclic_stream():
    """Eight interior children whose child pointers all lead back to their own sibling block: not a tree."""
    import struct
    b = bytes([1]) + struct.pack(">i", 7) + struct.pack(">H", 0)
    for i in range(8):
        b += bytes([1]) + struct.pack(">i", -7 * i) + struct.pack(">H", 0)
    return np.frombuffer(b, np.uint8).copy()


def test_cyclic_stream_is_refused(svo):
    import ctypes as C
    nodes = cyclic_stream()
    out = (C.c_uint64 * 8)()
    rc = svo._lib.lib().svo_transcode_probe(nodes.ctypes.data_as(C.c_void_p), nodes.size, 2, out, None, 0)
    assert rc == svo._lib.ERR_FORMAT


def _python_transcode(nodes):
    """Independent restatement of the transcode in Python (breadth-first, struct-based decoding)."""
    import struct
    b = bytes(nodes) + b"\0" * 16
    n = len(nodes)
    rd = lambda p: b[p] if p < n else 0
    be32 = lambda p: (rd(p) << 24) | (rd(p + 1) << 16) | (rd(p + 2) << 8) | rd(p + 3)
    be16 = lambda p: (rd(p) << 8) | rd(p + 1)
    cur = [(0, be32(1), be16(5))]
    out, base = [], 0
    for depth in range(23):
        if not cur:
            break
        nxt = []
        for off, cp, codes in cur:
            ref = (off + cp) & 0xFFFFFFFF
            p, nz, hd, first = ref, 0, 0, base + len(cur) + len(nxt)
            for c in range(8):
                code = (codes >> (2 * c)) & 3
                if rd(p):
                    nz |= 1 << c
                    ccp = be32(p + 1) if code == 0 else 0
                    if ccp and depth < 22:
                        hd |= 1 << c
                        nxt.append((p, ccp, be16(p + 5)))
                p = (p + (3 if code == 1 else 1 if code == 3 else 7)) & 0xFFFFFFFF
            out.append((first, codes | (nz << 16) | (hd << 24), ref))
        base += len(cur)
        cur = nxt
    return np.array(out, dtype=np.uint64)


def test_transcode_against_python_restatement(svo, oracle):
    rng = np.random.default_rng(21)
    for n in (8, 16, 32):
        vox = (rng.random((n, n, n)) < 0.25).astype(np.uint8) * rng.integers(1, 4, (n, n, n)).astype(np.uint8)
        vox[: n // 2, : n // 4] = 2
        nodes, _ = oracle.build_dense(vox)
        want = _python_transcode(nodes)
        out, desc = probe(svo, nodes, 3, want_desc=len(want))
        assert out[0] == len(want)
        assert np.array_equal(desc.astype(np.uint64), want)
