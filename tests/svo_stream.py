"""Test helper: hand-assemble and decode node streams in the reference layout
(/root/reference/src/engine/Octree.java:68-95,119-176), independently of the
oracle's C builder and of the product.

  interior          [value u8][childPtr i32 big-endian, relative to this node][leafMask u16 big-endian]   code 0
  surface leaf      [value u8][normal u16 little-endian]                                                  code 1
  subdividable leaf [value u8][6 x 0]                                                                     code 2
  non-surface leaf  [value u8]                                                                            code 3
Children are 8 consecutive records in child order n = x | y<<1 | z<<2.
"""
from __future__ import annotations

import struct

import numpy as np

SIZES = {0: 7, 1: 3, 2: 7, 3: 1}


class Node:
    """A node to serialise: kind in {"interior","surface","subdiv","nonsurf"}."""

    def __init__(self, kind, value=0, children=None, normal=0):
        self.kind, self.value, self.children, self.normal = kind, value, children, normal
        self.offset = None

    @property
    def code(self):
        return {"interior": 0, "surface": 1, "subdiv": 2, "nonsurf": 3}[self.kind]


def interior(value, children):
    assert len(children) == 8
    return Node("interior", value, children)


def surface(value, normal=555):
    return Node("surface", value, normal=normal)


def subdiv(value):
    return Node("subdiv", value)


def nonsurf(value):
    return Node("nonsurf", value)


def serialise(root: Node) -> np.ndarray:
    """Depth-first like Octree.constructInnerOctree: a node's 8 children are contiguous."""
    buf = bytearray()

    def put(n: Node):
        n.offset = len(buf)
        if n.kind in ("interior", "subdiv"):
            buf.extend(bytes([n.value]) + b"\0" * 6)
        elif n.kind == "surface":
            buf.extend(bytes([n.value]) + struct.pack("<H", n.normal))
        else:
            buf.extend(bytes([n.value]))

    def emit_children(n: Node):
        mask = 0
        for i, c in enumerate(n.children):
            put(c)
            mask |= c.code << (2 * i)
        rel = n.children[0].offset - n.offset
        buf[n.offset + 1:n.offset + 5] = struct.pack(">i", rel)
        buf[n.offset + 5:n.offset + 7] = struct.pack(">H", mask)
        for c in n.children:
            if c.kind == "interior" and c.children is not None:
                emit_children(c)

    put(root)
    emit_children(root)
    return np.frombuffer(bytes(buf), dtype=np.uint8).copy()


def tube(depth: int) -> Node:
    """Only the row of cells along x at y = z = 0 is subdivided, down to `depth`; every leaf is empty.  A ray sent
    along the row walks ~2.9 iterations per cell without ever hitting: the way to reach the 1500-iteration cap
    (svotrace.comp:264-266) and its neighbourhood with ordinary rays."""
    def build(d, iy, iz):
        kids = []
        for n in range(8):
            cy, cz = 2 * iy + ((n >> 1) & 1), 2 * iz + ((n >> 2) & 1)
            kids.append(build(d + 1, cy, cz) if (cy == 0 and cz == 0 and d + 1 < depth) else nonsurf(0))
        return interior(1, kids)
    return build(0, 0, 0)


def tube_rays(depth: int, n: int, seed: int = 7) -> np.ndarray:
    """Rays inside the tube, nearly parallel to it, from random start points (so their iteration counts cover a range)."""
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=np.dtype([("o", np.float32, 3), ("d", np.float32, 3)]))
    h = 2.0 ** -depth
    rays["o"][:, 0] = rng.uniform(1.0, 2.0, n)
    rays["o"][:, 1] = 1 + h * rng.uniform(0.2, 0.8, n)
    rays["o"][:, 2] = 1 + h * rng.uniform(0.2, 0.8, n)
    rays["d"][:, 0] = np.where(rng.random(n) < 0.5, -1, 1)
    rays["d"][:, 1] = rng.uniform(-1e-4, 1e-4, n)
    rays["d"][:, 2] = rng.uniform(-1e-4, 1e-4, n)
    return rays


def decode_voxels(nodes: np.ndarray, n: int) -> np.ndarray:
    """Walk a stream and paint every leaf into an n^3 grid [z,y,x] (interior nodes without children paint their value)."""
    out = np.zeros((n, n, n), np.uint8)
    b = nodes.tobytes()

    def walk(off, x, y, z, size):
        rel = struct.unpack(">i", b[off + 1:off + 5])[0]
        mask = struct.unpack(">H", b[off + 5:off + 7])[0]
        p = off + rel
        half = size // 2
        for i in range(8):
            code = (mask >> (2 * i)) & 3
            cx, cy, cz = x + (i & 1) * half, y + ((i >> 1) & 1) * half, z + ((i >> 2) & 1) * half
            val = b[p]
            if code == 0 and struct.unpack(">i", b[p + 1:p + 5])[0] != 0 and half > 1:
                walk(p, cx, cy, cz, half)
            else:
                out[cz:cz + half, cy:cy + half, cx:cx + half] = val
            p += SIZES[code]

    walk(0, 0, 0, 0, n)
    return out


def slab_hit(o, d, lo, hi):
    """float64 ray/box entry distance (None if missed)."""
    o, d, lo, hi = (np.asarray(v, np.float64) for v in (o, d, lo, hi))
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (lo - o) / d, (hi - o) / d
    tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
    tn = np.where(np.isnan(tn), -np.inf, tn)
    tf = np.where(np.isnan(tf), np.inf, tf)
    t_in, t_out = max(tn.max(), 0.0), tf.min()
    return float(t_in) if t_in <= t_out else None
