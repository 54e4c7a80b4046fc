"""A second, independent restatement of the reference traversal -- TEST INFRASTRUCTURE.

`intersect_octree` below was written from the GLSL text of /root/reference/src/shaders/svotrace.comp (getByte :75-79,
extractNode / extractLeaf / extractNonSurfaceLeaf / extractSubdividableLeaf :88-130, extractChild :132-157,
intersectOctree :211-432), statement by statement, in Python with numpy binary32 scalars -- without looking at
oracle/svo_oracle.c.  The reference has no golden vectors for this path and neither Java nor GLSL can run in the build
image, so the C oracle cannot be pinned against upstream outputs; what can be done is to have two independently written
restatements and demand that they agree bit for bit (tests/test_oracle_second_opinion.py).  Where GLSL leaves behaviour
to the driver both follow the arithmetic contract of DESIGN.md section 2: every operation one binary32 rounding, min /
max return the operand that is not a NaN, reads outside the buffer return 0.
"""
from __future__ import annotations

import numpy as np

F = np.float32
EPSILON = F(3.552713678800501e-15)
NODE_SIZE, LEAF_SIZE, NON_SURFACE_LEAF_SIZE = 7, 3, 1
MAX_SCALE = 23
MAX_RAYCAST_ITERATIONS = 1500
U32 = 0xFFFFFFFF


def _min(a, b):
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a < b else b


def _max(a, b):
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a > b else b


def _sign(x):
    return F(1.0) if x > 0 else (F(-1.0) if x < 0 else F(0.0))


def _bits(x) -> int:
    return int(np.array([x], dtype=np.float32).view(np.uint32)[0])


def _from_bits(u: int):
    return np.array([u & U32], dtype=np.uint32).view(np.float32)[0]


def _find_msb(x: int) -> int:
    return x.bit_length() - 1  # -1 for 0, like GLSL findMSB


class Node:
    __slots__ = ("value", "cp", "leafMask", "descriptor")

    def __init__(self, value, cp, leafMask, descriptor):
        self.value, self.cp, self.leafMask, self.descriptor = value, cp, leafMask, descriptor


class Buffer:
    """octreeBuffer (SSBO 7) as the shader sees it: a uint[] view of the byte stream; words past the end read 0."""

    def __init__(self, nodes: np.ndarray):
        raw = np.ascontiguousarray(nodes, dtype=np.uint8)
        pad = (-raw.size) % 4
        self.words = np.concatenate([raw, np.zeros(pad, np.uint8)]).view("<u4")

    def get_byte(self, byte_pointer: int) -> int:  # :75-79
        byte_pointer &= U32
        r = byte_pointer % 4
        i = byte_pointer // 4
        word = int(self.words[i]) if i < self.words.size else 0
        return (word & (0x000000FF << (r * 8))) >> (r * 8)

    def extract_node(self, p: int) -> Node:  # :88-101 (operands evaluated left to right)
        descriptor = p
        value = self.get_byte(p)
        cp = (self.get_byte(p + 1) << 24) | (self.get_byte(p + 2) << 16) | (self.get_byte(p + 3) << 8) | self.get_byte(p + 4)
        if cp & 0x80000000:
            cp -= 1 << 32  # int
        leaf_mask = (self.get_byte(p + 5) << 8) | self.get_byte(p + 6)
        return Node(value, cp, leaf_mask, descriptor)

    def extract_leaf(self, p: int) -> Node:  # :103-108
        n = Node(self.get_byte(p), 0, 0, p)
        n.leafMask = self.get_byte(p + 1) | (self.get_byte(p + 2) << 8)
        return n

    def extract_non_surface_leaf(self, p: int) -> Node:  # :110-114
        return Node(self.get_byte(p), 0, 0, p)

    def extract_subdividable_leaf(self, p: int) -> Node:  # :116-130
        n = self.extract_node(p)
        n.cp = 0
        return n

    def extract_child(self, parent_pointer: int, child_pointer: int, child: int, leaf_mask: int):  # :132-157 -> (Node, endPointer)
        i = 0
        pointer = ((child_pointer & U32) + parent_pointer) & U32
        while i < child:
            local = (leaf_mask & (0x0003 << (i << 1))) >> (i << 1)
            if local == 0 or local == 2:
                pointer += NODE_SIZE
            elif local == 1:
                pointer += LEAF_SIZE
            elif local == 3:
                pointer += NON_SURFACE_LEAF_SIZE
            i += 1
        pointer &= U32
        local = (leaf_mask & (0x0003 << (child << 1))) >> (child << 1)
        if local == 0:
            return self.extract_node(pointer), pointer
        if local == 1:
            return self.extract_leaf(pointer), pointer
        if local == 2:
            return self.extract_subdividable_leaf(pointer), pointer
        return self.extract_non_surface_leaf(pointer), pointer


def new_cast_result():
    """castResult (:186-197).  Uninitialised upstream; zero by the contract of DESIGN.md (U2)."""
    z = F(0.0)
    return {"value": 0, "pointer": 0, "iter": 0, "t": z, "scale": z, "debugColor": (z, z, z), "normal": (z, z, z), "voxelPos": (z, z, z), "depth": 0}


def intersect_octree(buf: Buffer, origin, direction, max_depth: int, cone_trace: bool, res: dict | None = None):
    """:211-432.  Returns a dict: hit, iter (loop iterations run, 1501 = capped) and, when the code after the loop ran, the
    castResult fields it writes.  `res`: the caller's castResult, updated the way the `out` parameter is (fields the
    shader does not write keep their old values)."""
    out = _intersect_octree(buf, origin, direction, max_depth, cone_trace)
    if res is not None:
        res["pointer"] = out.get("last_pointer", res["pointer"])  # extractChild writes res.pointer on every call (:294, :381)
        if out["capped"]:
            res["debugColor"] = (F(0.3), F(0.3), F(0.6))  # :213 survives
        elif "t" not in out:
            g = F(0.01) * F(out["iter"])  # :376
            res["debugColor"] = (g, g, g)
        else:
            for k in ("t", "value", "normal", "scale", "depth", "voxelPos"):
                res[k] = out[k]
            res["iter"] = out["iter"]
            g = F(0.005) * F(out["iter"])  # :428
            res["debugColor"] = (g, g, g)
    return out


def _intersect_octree(buf: Buffer, origin, direction, max_depth: int, cone_trace: bool):
    ox, oy, oz = (F(v) for v in origin)
    dx, dy, dz = (F(v) for v in direction)
    with np.errstate(all="ignore"):
        parent = buf.extract_node(0)
        it = 0
        if abs(dx) < EPSILON:
            dx = EPSILON * _sign(dx)
        if abs(dy) < EPSILON:
            dy = EPSILON * _sign(dy)
        if abs(dz) < EPSILON:
            dz = EPSILON * _sign(dz)
        tx_coef = F(1.0) / -abs(dx)
        ty_coef = F(1.0) / -abs(dy)
        tz_coef = F(1.0) / -abs(dz)
        tx_bias = tx_coef * ox
        ty_bias = ty_coef * oy
        tz_bias = tz_coef * oz
        octant_mask = 0
        if dx > 0:
            octant_mask ^= 1
            tx_bias = F(3.0) * tx_coef - tx_bias
        if dy > 0:
            octant_mask ^= 2
            ty_bias = F(3.0) * ty_coef - ty_bias
        if dz > 0:
            octant_mask ^= 4
            tz_bias = F(3.0) * tz_coef - tz_bias
        t_min = _max(_max(F(2.0) * tx_coef - tx_bias, F(2.0) * ty_coef - ty_bias), F(2.0) * tz_coef - tz_bias)
        t_max = _min(_min(tx_coef - tx_bias, ty_coef - ty_bias), tz_coef - tz_bias)
        t_min = _max(t_min, F(0.0))
        h = t_max
        idx = 0
        px, py, pz = F(1.0), F(1.0), F(1.0)
        scale = MAX_SCALE - 1
        scale_exp2 = F(0.5)
        child_descriptor = 0
        if F(1.5) * tx_coef - tx_bias > t_min:
            idx ^= 1
            px = F(1.5)
        if F(1.5) * ty_coef - ty_bias > t_min:
            idx ^= 2
            py = F(1.5)
        if F(1.5) * tz_coef - tz_bias > t_min:
            idx ^= 4
            pz = F(1.5)
        child_shift = 0
        stack = [None] * (MAX_SCALE + 1)
        pointer = 0
        stale_pops = 0
        while scale < MAX_SCALE:
            it += 1
            if it > MAX_RAYCAST_ITERATIONS:
                return {"hit": False, "iter": it, "capped": True, "stale_pops": stale_pops, "last_pointer": pointer}
            if child_descriptor == 0:
                child_descriptor = parent.cp
            if t_min > F(0.05) and cone_trace:
                max_depth = 11
            tx_corner = px * tx_coef - tx_bias
            ty_corner = py * ty_coef - ty_bias
            tz_corner = pz * tz_coef - tz_bias
            tc_max = _min(_min(tx_corner, ty_corner), tz_corner)
            child_shift = idx ^ octant_mask
            child, pointer = buf.extract_child(parent.descriptor, child_descriptor, child_shift, parent.leafMask)
            if child.value != 0 and t_min <= t_max:
                if MAX_SCALE - scale == max_depth:
                    break
                tv_max = _min(t_max, tc_max)
                one_half = scale_exp2 * F(0.5)
                tx_center = one_half * tx_coef + tx_corner
                ty_center = one_half * ty_coef + ty_corner
                tz_center = one_half * tz_coef + tz_corner
                if t_min <= tv_max:
                    if child.cp == 0:
                        break
                    if tc_max < h:
                        stack[scale] = (parent, t_max)
                    h = tc_max
                    parent = child
                    idx = 0
                    scale -= 1
                    scale_exp2 = one_half
                    if tx_center > t_min:
                        idx ^= 1
                        px = px + scale_exp2
                    if ty_center > t_min:
                        idx ^= 2
                        py = py + scale_exp2
                    if tz_center > t_min:
                        idx ^= 4
                        pz = pz + scale_exp2
                    t_max = tv_max
                    child_descriptor = 0
                    continue
            step_mask = 0
            if tx_corner <= tc_max:
                step_mask ^= 1
                px = px - scale_exp2
            if ty_corner <= tc_max:
                step_mask ^= 2
                py = py - scale_exp2
            if tz_corner <= tc_max:
                step_mask ^= 4
                pz = pz - scale_exp2
            t_min = tc_max
            idx ^= step_mask
            if (idx & step_mask) != 0:
                differing_bits = 0
                if step_mask & 1:
                    differing_bits |= _bits(px) ^ _bits(px + scale_exp2)
                if step_mask & 2:
                    differing_bits |= _bits(py) ^ _bits(py + scale_exp2)
                if step_mask & 4:
                    differing_bits |= _bits(pz) ^ _bits(pz + scale_exp2)
                scale = _find_msb(differing_bits)
                scale_exp2 = _from_bits(((scale - MAX_SCALE + 127) & U32) << 23)
                if not (0 <= scale <= MAX_SCALE) or stack[scale] is None:
                    # Index 23 (the ray left the cube) is never written upstream either; the loop condition ends the cast
                    # right after, so the values do not matter.  Any other unwritten entry would be a read of what an
                    # earlier cast left in the global array: reported, the tests demand that it never happens.
                    if scale < MAX_SCALE:
                        stale_pops += 1
                    parent, t_max = Node(0, 0, 0, 0), F(0.0)
                else:
                    parent, t_max = stack[scale]
                shx = _bits(px) >> scale
                shy = _bits(py) >> scale
                shz = _bits(pz) >> scale
                px = _from_bits(shx << scale)
                py = _from_bits(shy << scale)
                pz = _from_bits(shz << scale)
                idx = (shx & 1) | ((shy & 1) << 1) | ((shz & 1) << 2)
                h = F(0.0)
                child_descriptor = 0
        if scale >= MAX_SCALE:
            return {"hit": False, "iter": it, "capped": False, "stale_pops": stale_pops, "last_pointer": pointer}
        nx, ny, nz = F(0.0), F(0.0), F(0.0)
        target, pointer = buf.extract_child(parent.descriptor, child_descriptor, child_shift, parent.leafMask)
        if target.leafMask != 0:
            raw = target.leafMask
            vx = F((raw % 10) - 5)
            vy = F((((raw % 100) - (raw % 10)) // 10) - 5)
            vz = F(((raw - (raw % 100)) // 100) - 5)
            length = np.sqrt((vx * vx + vy * vy) + vz * vz)  # normalize = v / sqrt(dot(v, v)), contract of DESIGN.md
            nx, ny, nz = vx / length, vy / length, vz / length
        vpx, vpy, vpz = px, py, pz
        if dx > 0:
            vpx = F(3.0) - vpx - scale_exp2
        if dy > 0:
            vpy = F(3.0) - vpy - scale_exp2
        if dz > 0:
            vpz = F(3.0) - vpz - scale_exp2
        vpx = vpx + nx * scale_exp2 * F(2.0) * F(1.74)
        vpy = vpy + ny * scale_exp2 * F(2.0) * F(1.74)
        vpz = vpz + nz * scale_exp2 * F(2.0) * F(1.74)
        return {"hit": bool(scale < MAX_SCALE and t_min <= t_max), "iter": it, "capped": False, "stale_pops": stale_pops, "pointer": pointer,
                "last_pointer": pointer, "t": t_min,
                "value": target.value, "normal": (nx, ny, nz), "scale": scale_exp2, "depth": MAX_SCALE - scale,
                "voxelPos": (vpx, vpy, vpz)}


# ---------------------------------------------------------------------------------------------------------------
# trace() :435-646 and main() :649-729.  sin / cos / acos / exp are the contract's fixed kernels (DESIGN.md section 2;
# tested against libm on their own): the caller passes them in, everything else is restated here.
# ---------------------------------------------------------------------------------------------------------------
PI = F(3.14159265359)
SQRT3 = F(1.73205080757)
MAX_DEPTH = 13


def _v(a, b, c):
    return (F(a), F(b), F(c))


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _normalize(v):
    length = np.sqrt(_dot(v, v))
    return (v[0] / length, v[1] / length, v[2] / length)


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _mix(x, y, a):
    return x * (F(1.0) - a) + y * a


def _rand(m, x, y):  # :26-29
    s = m["sin"](x * F(12.9898) + y * F(78.233))
    t = s * F(43758.5453)
    return t - np.floor(t)


def trace(buf, m, beam_dist, origin, direction, seed0, seed1, seed2, render_mode, max_depth=MAX_DEPTH, casts=2):
    """Returns (colour, depth or None if never written, primary: dict of what the validation planes record)."""
    with np.errstate(all="ignore"):
        res = new_cast_result()
        res["t"] = F(2.0)
        origin = tuple(origin[k] + direction[k] * beam_dist for k in range(3))
        d = direction
        depth = None
        accum = _v(0, 0, 0)
        mask = _v(1, 1, 1)
        primary = None

        def cast(o, dd, cone):
            nonlocal primary
            out = intersect_octree(buf, o, dd, max_depth, cone, res)
            if primary is None:
                primary = {"hit": out["hit"], "iter": out["iter"], "pointer": out.get("pointer"), "t": out.get("t")}
            return out["hit"]

        if render_mode == 0:
            for i in range(casts):
                intersect = cast(origin, d, i != 0)
                if not intersect and i == 0:
                    sky = _v(0.6725, 0.8784, 1.0)
                    k = _v(0.4, 0.4, 0.25)
                    accum = tuple(accum[c] + (sky[c] - d[1] * k[c]) for c in range(3))
                    break
                normal = res["normal"]
                hitpoint = res["voxelPos"]
                r = _rand(m, seed0 + _rand(m, seed0, seed2 * F(0.1)), seed1 + _rand(m, seed1, seed2 * F(0.02)))
                rand1 = F(2.0) * PI * r
                w = normal
                axis = _v(0, 1, 0) if abs(w[0]) > F(0.1) else _v(1, 0, 0)
                u = _normalize(_cross(axis, w))
                v = _cross(w, u)
                c1, s1, omr = m["cos"](rand1), m["sin"](rand1), F(1.0) - r
                newdir = _normalize(tuple((u[c] * c1 + v[c] * s1) + w[c] * omr for c in range(3)))
                origin = hitpoint
                d = newdir
                matcolor = tuple(hitpoint[c] - F(1.0) for c in range(3))
                if res["value"] == 1:
                    matcolor = _v(0.84, 0.86, 0.78)
                if res["value"] == 2:
                    matcolor = _v(0.57, 0.5, 0.31)
                if res["value"] == 3:
                    matcolor = _v(0.37, 0.43, 0.27)
                if intersect:
                    depth = res["t"]
                    accum = tuple(accum[c] + mask[c] * F(0.0) for c in range(3))
                    dn = _dot(newdir, normal)
                    mask = tuple(mask[c] * matcolor[c] * dn for c in range(3))
                else:
                    sun = _normalize(_v(1, 1, 1))
                    diff = m["acos"](_dot(d, sun))
                    if diff < F(0.4):
                        accum = tuple(accum[c] + mask[c] * F(7.0) for c in range(3))
                    accum = tuple(accum[c] + mask[c] * F(1.0) for c in range(3))
                    depth = F(0.0)
                    break
            return accum, depth, primary
        if render_mode == 1:
            if cast(origin, d, False):
                depth = res["t"]
            else:
                depth = F(0.0)
            return res["debugColor"], depth, primary
        if render_mode == 2:
            if cast(origin, d, False):
                depth = res["t"]
                matcolor = _v(0, 0, 0)
                if res["value"] == 1:
                    matcolor = _v(0.84, 0.86, 0.78)
                if res["value"] == 2:
                    matcolor = _v(0.57, 0.5, 0.31)
                if res["value"] == 3:
                    matcolor = _v(0.37, 0.43, 0.27)
                sun = _normalize(_v(0.5, 0.5, 0.5))
                if res["depth"] >= 10:
                    ph = _dot(res["normal"], sun) * F(0.1)
                else:
                    ph = _dot(_v(0, 1, 0), sun) * F(0.1)
                matcolor = tuple(matcolor[c] + ph for c in range(3))
                true_dist = res["t"] + beam_dist
                lg = m["exp"](F(-0.5) * true_dist * F(2.0))
                lb = m["exp"](F(-0.5) * true_dist * F(4.0))
                lr = m["exp"](F(-0.5) * true_dist * F(1.0))
                matcolor = (lr * matcolor[0] + (F(1.0) - lr) * F(1.0), lg * matcolor[1] + (F(1.0) - lg) * F(1.0),
                            lb * matcolor[2] + (F(1.0) - lb) * F(1.0))
                if cast(res["voxelPos"], sun, False) and res["t"] > res["scale"] * SQRT3:
                    matcolor = tuple(matcolor[c] - F(0.2) for c in range(3))
                elif res["iter"] > 260:
                    pen = F(0.05) * F(res["iter"]) / F(100.0)
                    matcolor = tuple(matcolor[c] - pen for c in range(3))
                return matcolor, depth, primary
            sky = _v(0.6725, 0.8784, 1.0)
            k = _v(0.4, 0.4, 0.25)
            return tuple(sky[c] - d[1] * k[c] for c in range(3)), F(0.0), primary
        if render_mode == 3:
            if cast(origin, d, False):
                return tuple(res["normal"][c] * F(0.5) + F(0.5) for c in range(3)), res["t"], primary
            return _v(0, 0, 0), F(0.0), primary
        return res["voxelPos"], depth, primary  # mode 4


def render_pixel(buf, m, cam_pos, l1, l2, r1, r2, frame_number, render_mode, width, height, x, y, max_depth=MAX_DEPTH, casts=2):
    """main() for one pixel (no beam pass).  Returns (colour after the debug overlay, depth, primary)."""
    with np.errstate(all="ignore"):
        px = (F(x) + F(0.5)) / F(width)
        py = (F(y) + F(0.5)) / F(height)
        d = tuple(_mix(_mix(F(l1[k]), F(l2[k]), py), _mix(F(r1[k]), F(r2[k]), py), px) for k in range(3))
        nd = _normalize(d)
        color, depth, primary = trace(buf, m, F(0.0), tuple(F(v) for v in cam_pos), nd, F(x), F(y), F(frame_number), render_mode, max_depth, casts)
        if depth is None:
            depth = F(-1.0)  # :672
        if x < 10 and y < 10:
            first_word = int(buf.words[0]) if buf.words.size else 0
            color = _v(1, 0, 0) if first_word == 0 else _v(1, 1, 1)
        return color, depth, primary
