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


class StreamEditor:
    """In-place edits of a node stream the way the engine's SDF brush makes them (Octree.java:700-885): values set in
    existing records, subdividable leaves (7-byte records) turned into interior nodes whose eight children are APPENDED at
    memOffset, interior children collapsed back into leaves.  Tracks the two byte ranges the engine would push with
    Renderer.updateSSBO (ChangeBounds, Octree.java:676-698): [start0, end0) over touched existing records and
    [start1, end1) over appended ones."""

    def __init__(self, nodes: np.ndarray, slack: int = 1 << 16):
        self.buf = bytearray(nodes.tobytes()) + bytearray(slack)
        self.mem = int(nodes.size)  # memOffset
        self.start0, self.end0 = self.mem, 0
        self.start1 = self.end1 = self.mem

    # -- record access -------------------------------------------------------------------------------------------
    def cp(self, off):
        return struct.unpack(">i", self.buf[off + 1:off + 5])[0]

    def mask(self, off):
        return struct.unpack(">H", self.buf[off + 5:off + 7])[0]

    def child(self, parent_off, n):
        """(record offset, type code) of child n of the interior record at parent_off."""
        m, p = self.mask(parent_off), parent_off + self.cp(parent_off)
        for i in range(n):
            p += SIZES[(m >> (2 * i)) & 3]
        return p, (m >> (2 * n)) & 3

    def walk(self, path):
        """Follow child indices from the root.  Returns (parent offset, child index, record offset, code) of the last step."""
        off, parent, code = 0, None, 0
        for n in path:
            assert code == 0 and self.cp(off) != 0, "path leaves the interior nodes"
            parent = off
            off, code = self.child(off, n)
        return parent, path[-1], off, code

    def _touch(self, a, b):  # updateExistingNodeBounds
        if a < self.start1:
            self.start0 = min(self.start0, a)
            self.end0 = max(self.end0, min(b, self.start1))

    # -- edits ---------------------------------------------------------------------------------------------------
    def set_value(self, path, value):
        _, _, off, _ = self.walk(path)
        self.buf[off] = value
        self._touch(off, off + 7)

    def set_code(self, parent_off, n, code):
        m = self.mask(parent_off)
        m = (m & ~(3 << (2 * n))) | (code << (2 * n))
        self.buf[parent_off + 5:parent_off + 7] = struct.pack(">H", m)
        self._touch(parent_off, parent_off + 7)

    def subdivide(self, path, values, surface=False, normal=595):
        """subdivideNode (Octree.java:829-885): the 7-byte leaf at `path` becomes an interior node with eight appended
        children (subdividable leaves, or 3-byte surface leaves when `surface`) holding `values`."""
        parent, n, off, code = self.walk(path)
        assert code == 2, "only subdividable leaves have room for a child pointer"
        self.set_code(parent, n, 0)
        first, m = self.mem, 0
        for i, v in enumerate(values):
            if surface:
                self.buf[self.mem:self.mem + 3] = bytes([v]) + struct.pack("<H", normal)
                self.mem += 3
                m |= 1 << (2 * i)
            else:
                self.buf[self.mem:self.mem + 7] = bytes([v]) + b"\0" * 6
                self.mem += 7
                m |= 2 << (2 * i)
        self.buf[off + 1:off + 5] = struct.pack(">i", first - off)
        self.buf[off + 5:off + 7] = struct.pack(">H", m)
        if any(values) and self.buf[off] == 0:
            self.buf[off] = max(values)
        self._touch(off, off + 7)
        self.end1 = self.mem

    def collapse(self, path, value):
        """An interior child becomes a subdividable leaf again (the 'fully inside the volume' case, Octree.java:770-784)."""
        parent, n, off, code = self.walk(path)
        assert code == 0
        self.buf[off] = value
        self.set_code(parent, n, 2)
        self._touch(off, off + 7)

    def stream(self) -> np.ndarray:
        return np.frombuffer(bytes(self.buf[:self.mem]), dtype=np.uint8).copy()

    def ranges(self):
        r = []
        if self.end0 > self.start0:
            r.append((self.start0, self.end0))
        if self.end1 > self.start1:
            r.append((self.start1, self.end1))
        return r

    def find_leaf(self, want_code, want_nonzero, rng, min_depth=2, tries=4000):
        """A random path to a record of the given type (value zero / non-zero as asked)."""
        for _ in range(tries):
            off, path, code = 0, [], 0
            while True:
                if code != 0 or self.cp(off) == 0:
                    break
                n = int(rng.integers(0, 8))
                off, code = self.child(off, n)
                path.append(n)
                if len(path) > 20:
                    break
            if code == want_code and (self.buf[off] != 0) == want_nonzero and len(path) >= min_depth:
                return path
        raise AssertionError("no such record found")
