#!/usr/bin/env python3
"""Builds oracle/_ref/libsvo_ref_java.so: THE REFERENCE'S OWN OCTREE BUILDER AND SDF BRUSH, compiled for the CPU.
TEST INFRASTRUCTURE.

The reference's world builder and edit path are Java (src/engine/Octree.java, OctreeThread.java, Util.java,
Constants.java, sdf/*.java) and no JVM exists in this image.  The methods on this path are, however, C-like text: integer
arithmetic, loops, absolute ByteBuffer puts.  This script reads them where they lie under /root/reference (nothing is
copied into the repository: the generated C++ goes to oracle/_ref/, git-ignored), applies the MECHANICAL rewrite below --
syntax only -- and compiles the result with g++ against oracle/java_shim.h (the slice of the Java runtime they touch) and
oracle/ref_java_harness.cpp (extern "C" entry points; what Main / WorldGenerator do around the builder).  Control flow,
scan order, record layout, normals, exposure tests, the chunk splice and the brush's case analysis are the reference's,
parsed by a C++ compiler.

What is taken (member names; everything else of the classes -- GL, PNG and file I/O, printing -- is left out):
  Octree: fields buffer / memOffset / bufferSize / the four counters / NODE_SIZE / LEAF_SIZE / NON_SURFACE_LEAF_SIZE /
    CHUNK_SIZE / marchTime / childOffsets; classes Chunk, NormalResult, ChangeBounds, NodeInfo; enum NodeType; the
    constructor; createDummyHead, getValue, setValue, getVoxel, setNormal, the four create*Node, set/getChildPointer,
    set/getLeafMask, fillEmptyChildren, constructInnerOctree, genSurfaceNormal, checkBigNodeExposed,
    updateExistingNodeBounds, useSDFBrush (both), subdivideNode, forEachChild (both), genChildPositions, markNodeAsDirty;
    and lines `int[] startPos = ...` to the end of the splice loop of constructCompleteOctree(shader, voxelTexture,
    heightmapTexture, materialTexture) (Octree.java:285-338) as the body of a method buildChunk(chunk, voxelBuffer, maxLOD).
  OctreeThread, Util, sdf/SignedDistanceField, sdf/Sphere, sdf/Box: whole classes.  Constants: the integer constants.
  src/shaders/chunkgen-heightmap.comp (the voxeliser the builder reads its voxels from): whole, through oracle/build_ref.py's
    rules R1-R3 and oracle/glsl_voxel_shim.h (generate_voxeliser below).

The rewrite, in full (regular expressions over the comment-stripped text unless noted):
  J1  `byte short long boolean` -> `jbyte jshort jlong bool`; `null` -> `nullptr`; `public private protected final
      abstract @Override` dropped; `static final T` -> `static inline T` (settable: J9); `class X extends Y` ->
      `struct X : Y`; every class body gets its `;`.
  J2  arrays are references: `T[]` -> `jarray<T>`, `T[][]` -> `jarray<jarray<T>>`; `new T[n]` -> `jarray<T>::make(n)`;
      `new T[a][b]` -> `make2<T>(a, b)`; `new T[] {..}` and `= {..}` initialisers -> `jarray<T>{..}`.
  J3  objects are pointers: a variable, parameter, field or lambda parameter of class type C is declared `C*`; `x.` becomes
      `x->` for the names so declared (listed in OBJECTS below), for `this`, for `threads[i]` and after a call that returns
      a buffer (`).put(` `).limit(`); `new C(..)` stays `new C(..)`; `List<Chunk>` / `ArrayList<Chunk>` ->
      `ArrayList<Chunk*>*`.
  J4  statics: `Math. Util. Constants. System. BufferUtils.` -> `::`; `ByteOrder.BIG_ENDIAN` -> `ByteOrder::BIG_ENDIAN_`
      (BIG_ENDIAN is a libc macro); `NodeType.X` -> `X`.
  J5  lambdas: `(info) -> expr;` -> `[&](NodeInfo* info) { expr; };`, `(info) -> { .. };` -> `[&](NodeInfo* info) { .. };`;
      `method.accept(x)` -> `method(x)`.
  J6  `(int) <primary>` -> `J::to_int(<primary>)`: Java's saturating, NaN -> 0 narrowing (Sphere.normal at the centre).
  J7  a `case X:` that ends a switch gets an empty statement (C++17 wants one); `System.out.println(..);` dropped.
  J8  the non-static inner class ChangeBounds reads the enclosing object's memOffset: its constructor takes the enclosing
      `Octree*` (`new ChangeBounds()` -> `new ChangeBounds(this)`), `memOffset` inside it -> `outer->memOffset`.
  J9  constants that the oracle generalises become variables WHOSE DEFAULTS ARE THE SHIPPED VALUES: CHUNK_SIZE (1024) and
      Constants.* through J1; OctreeThread.run's literal `512` -> `Constants::SUB_OCTREE_SIZE` (512); the public useSDFBrush's
      literal `13` -> `Constants::SDF_MAX_LOD` (13).  getVoxel's index `x | (y << 10) | (z << 20)` is untouched: the voxel
      buffer keeps the reference's 1024 pitch whatever the chunk size.
  J10 SignedDistanceField's methods become `virtual`.

Usage: python oracle/build_ref_java.py [--force]     (needs /root/reference; the GPU box never runs it)
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
REF = "/root/reference/src/engine"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libsvo_ref_java.so")
GEN = os.path.join(OUT, "ref_java_gen.inc")
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing", "-w"]

CLASSES = ["Chunk", "NormalResult", "ChangeBounds", "NodeInfo", "Octree", "OctreeThread", "SignedDistanceField", "Sphere", "Box", "ByteBuffer"]
# names declared with a class type somewhere in the text taken (J3)
OBJECTS = ["buffer", "voxelData", "voxelBuffer", "childBuffer", "chunk", "chunks", "changeBounds", "cb", "info", "thread",
           "normalResult", "sdf", "octree", "this"]

OCTREE_FIELDS = ["buffer", "memOffset", "bufferSize", "surfaceLeafNodes", "nonSurfaceLeafNodes", "interiorNodes",
                 "subdividableLeafNodes", "NODE_SIZE", "LEAF_SIZE", "NON_SURFACE_LEAF_SIZE", "CHUNK_SIZE", "marchTime", "childOffsets"]
OCTREE_TYPES = ["Chunk", "NormalResult", "ChangeBounds", "NodeInfo", "NodeType"]
OCTREE_METHODS = ["Octree", "createDummyHead", "getValue", "setValue", "getVoxel", "setNormal", "createInteriorNode",
                  "createSubdividableLeafNode", "createSurfaceLeafNode", "createNonSurfaceLeafNode", "setChildPointer",
                  "getChildPointer", "setLeafMask", "getLeafMask", "fillEmptyChildren", "constructInnerOctree", "genSurfaceNormal",
                  "checkBigNodeExposed", "updateExistingNodeBounds", "useSDFBrush", "subdivideNode", "forEachChild",
                  "genChildPositions", "markNodeAsDirty"]
CONSTANTS = ["OCTREE_MEMORY_SIZE_KB", "SUB_OCTREE_MEMORY_SIZE_KB", "CHUNK_SIZE", "DELETE_VALUE", "WORLD_SIZE", "MARCH_DISTANCE_MIN_CUTOFF"]


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def match_brace(src: str, open_at: int) -> int:
    """Index just past the `}` that closes the `{` at open_at."""
    depth = 0
    for i in range(open_at, len(src)):
        if src[i] == "{":
            depth += 1
        elif src[i] == "}":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced braces")


def class_body(src: str, name: str) -> str:
    m = re.search(r"\bclass\s+%s\b[^{]*\{" % name, src)
    return src[m.end():match_brace(src, m.end() - 1) - 1]


def members(body: str):
    """Top-level members of a class body: (header text up to `{` `=` or `;`, whole text)."""
    out, i, n = [], 0, len(body)
    while i < n:
        while i < n and body[i].isspace():
            i += 1
        if i >= n:
            break
        j, depth = i, 0
        while j < n:
            c = body[j]
            if c == "{":
                if depth == 0 and "=" not in body[i:j]:  # a method / class / enum body: ends at its `}`
                    j = match_brace(body, j)
                    break
                depth += 1
            elif c == "}":
                depth -= 1
            elif c == ";" and depth == 0:
                j += 1
                break
            j += 1
        out.append(body[i:j])
        i = j
    return out


def member_name(text: str) -> str:
    head = re.split(r"[{=;]", text, 1)[0]
    head = re.sub(r"\([^)]*\)?.*", "", head, flags=re.S)  # drop a parameter list
    words = re.findall(r"[A-Za-z_]\w*", head)
    return words[-1] if words else ""


def wrap_int_casts(t: str) -> str:                                                               # J6
    out, i = [], 0
    for m in re.finditer(r"\(int\)\s*", t):
        if m.start() < i:
            continue
        out.append(t[i:m.start()])
        j = m.end()
        k = j
        while k < len(t) and (t[k].isalnum() or t[k] in "_:.>-"):
            if t[k] == "-" and t[k:k + 2] != "->":
                break
            k += 1
        while k < len(t) and t[k] in "([":
            close = {"(": ")", "[": "]"}[t[k]]
            depth = 0
            while True:
                if t[k] in "([":
                    depth += 1
                elif t[k] in ")]":
                    depth -= 1
                k += 1
                if depth == 0:
                    break
            _ = close
        out.append("J::to_int(" + t[j:k] + ")")
        i = k
    out.append(t[i:])
    return "".join(out)


def rewrite(t: str, owner: str = "") -> str:
    t = re.sub(r"System\.out\.println\((?:[^;]|\n)*?\);", "", t)                                   # J7
    t = re.sub(r"@Override", "", t)                                                                # J1
    t = re.sub(r"\bstatic\s+final\s+", "static inline ", t)
    t = re.sub(r"\bpublic\s+static\s+final\s+", "static inline ", t)
    t = re.sub(r"\b(public|private|protected|final|abstract)\s+", "", t)
    t = re.sub(r"\bstatic\s+(?!inline)(\w+(?:\[\])*)\s+(\w+)\s*=", r"static inline \1 \2 =", t)      # static byte[][] childOffsets = ...
    t = re.sub(r"\bboolean\b", "bool", t)
    t = re.sub(r"\bbyte\b", "jbyte", t)
    t = re.sub(r"\bshort\b", "jshort", t)
    t = re.sub(r"\blong\b", "jlong", t)
    t = re.sub(r"\bnull\b", "nullptr", t)
    t = re.sub(r"\bclass\s+(\w+)\s+extends\s+(\w+)", r"struct \1 : \2", t)
    t = re.sub(r"\bclass\s+(\w+)", r"struct \1", t)
    t = re.sub(r"\benum\s+(\w+)", r"enum \1", t)
    # J5 lambdas (before J3 touches `info.`)
    t = re.sub(r"\((\w+)\)\s*->\s*\{", r"[&](NodeInfo* \1) {", t)
    t = re.sub(r"\((\w+)\)\s*->\s*((?:[^;{]|\n)*?);", r"[&](NodeInfo* \1) { \2; };", t)
    t = re.sub(r"\bmethod\.accept\(", "method(", t)
    t = re.sub(r"Consumer<NodeInfo>", "Consumer<NodeInfo*>", t)
    # J2 arrays
    t = re.sub(r"\bnew\s+(\w+)\s*\[([^\]]+)\]\s*\[([^\]]+)\]", r"make2<\1>(\2, \3)", t)
    t = re.sub(r"\bnew\s+(\w+)\s*\[\]\s*\{", r"jarray<\1>{", t)
    t = re.sub(r"\bnew\s+(\w+)\s*\[([^\]]+)\]", lambda m: "jarray<%s>::make(%s)" % (m.group(1) + ("*" if m.group(1) in CLASSES else ""), m.group(2)), t)
    t = re.sub(r"\b(\w+)\s*\[\]\s*\[\]\s+(\w+)\s*=\s*\{", r"jarray<jarray<\1>> \2 = jarray<jarray<\1>>{", t)
    t = re.sub(r"\b(\w+)\s*\[\]\s+(\w+)\s*=\s*\{", r"jarray<\1> \2 = jarray<\1>{", t)
    t = re.sub(r"\b(\w+)\s*\[\]\s*\[\]", r"jarray<jarray<\1>>", t)
    t = re.sub(r"\b(\w+)\s*\[\]", lambda m: "jarray<%s>" % (m.group(1) + ("*" if m.group(1) in CLASSES else "")), t)
    # J3 objects are pointers
    t = re.sub(r"\b(?:Array)?List<Chunk>\s+(\w+)\s*=\s*new\s+ArrayList<Chunk>\(\)", r"ArrayList<Chunk*>* \1 = new ArrayList<Chunk*>()", t)
    t = re.sub(r"\b(?:Array)?List<Chunk>", "ArrayList<Chunk*>*", t)
    for c in CLASSES:
        t = re.sub(r"(?<![\w<:*])%s\s+(\w+)\s*(?=[;=,):])" % c, r"%s* \1" % c, t)                  # declarations and parameters
        t = re.sub(r"(?<![\w<:*])%s\s+(\w+)\s*\(" % c, r"%s* \1(" % c, t)                           # return types
    t = re.sub(r"new\s+(\w+)\*", r"new \1", t)
    for o in OBJECTS:
        t = re.sub(r"\b%s\." % o, "%s->" % o, t)
    t = re.sub(r"\bthreads\[(\w+)\]\.", r"threads[\1]->", t)
    t = re.sub(r"\)\.(put|limit|position)\(", r")->\1(", t)
    # J4 statics
    t = re.sub(r"\bByteOrder\.BIG_ENDIAN\b", "ByteOrder::BIG_ENDIAN_", t)
    t = re.sub(r"\b(Math|Util|Constants|System|BufferUtils)\.", r"\1::", t)
    t = re.sub(r"\bNodeType\.", "", t)
    t = wrap_int_casts(t)                                                                          # J6
    t = re.sub(r"(case\s+\w+\s*:)(\s*\})", r"\1 ;\2", t)                                           # J7
    return t


def close_classes(t: str) -> str:
    """`struct X { .. }` / `enum X { .. }` -> with the `;` C++ wants (J1)."""
    out, i = [], 0
    for m in re.finditer(r"\b(?:struct|enum)\s+\w+[^{;()]*\{", t):
        if m.start() < i:
            continue
        end = match_brace(t, m.end() - 1)
        out.append(t[i:m.end()])
        out.append(close_classes(t[m.end():end - 1]))
        out.append("};")
        i = end
    out.append(t[i:])
    return "".join(out)


def generate() -> str:
    def read(rel):
        with open(os.path.join(REF, rel)) as f:
            return strip_comments(f.read())

    octree = read("Octree.java")
    body = class_body(octree, "Octree")
    taken, splice = [], None
    for mem in members(body):
        name = member_name(mem)
        if name in OCTREE_FIELDS and "(" not in re.split(r"[={]", mem, 1)[0]:
            taken.append(mem)
        elif name in OCTREE_TYPES or (name in OCTREE_METHODS and "(" in mem.split("{", 1)[0]):
            taken.append(mem)
        elif name == "constructCompleteOctree" and "heightmapTexture" in mem.split("{", 1)[0]:
            a = mem.index("int[] startPos")
            b = mem.index("subdividableLeafNodes += thread.octree.subdividableLeafNodes;")
            b = mem.index("}", b) + 1
            splice = "void buildChunk(Chunk chunk, ByteBuffer voxelBuffer, int maxLOD) {\n" + mem[a:b] + "\n}\n"
    assert splice is not None, "chunk splice not found"
    missing = [n for n in OCTREE_FIELDS + OCTREE_TYPES + OCTREE_METHODS if not any(member_name(m) == n for m in taken)]
    assert not missing, "members not found in Octree.java: %s" % missing
    # J8: the inner class that reads the enclosing object's field
    oct_text = "\n".join(taken) + "\nvoid buildChunk(Chunk chunk, ByteBuffer voxelBuffer, int maxLOD);\n"
    oct_text = re.sub(r"ChangeBounds\(\)\s*\{((?:[^{}]|\n)*?)\}",
                      lambda m: "ChangeBounds(Octree outer) {" + re.sub(r"\bmemOffset\b", "outer.memOffset", m.group(1)) + "}", oct_text, count=1)
    oct_text = oct_text.replace("new ChangeBounds()", "new ChangeBounds(this)")
    oct_text = re.sub(r"(useSDFBrush\(sdf, 0, 0, 0, Constants\.WORLD_SIZE, pos, false, value, 0, )13(, changeBounds\))", r"\1Constants.SDF_MAX_LOD\2", oct_text)  # J9
    assert "Constants.SDF_MAX_LOD" in oct_text
    oct_cpp = rewrite(oct_text).replace("outer.memOffset", "outer->memOffset")

    thread = read("OctreeThread.java")
    thread = thread[thread.index("public class OctreeThread"):]
    thread = re.sub(r"constructInnerOctree\(512,", "constructInnerOctree(Constants.SUB_OCTREE_SIZE,", thread)   # J9
    assert "Constants.SUB_OCTREE_SIZE" in thread

    util = read("Util.java")
    util = util[util.index("public class Util"):]
    sdfs = []
    for f in ("sdf/SignedDistanceField.java", "sdf/Sphere.java", "sdf/Box.java"):
        s = read(f)
        s = s[s.index("public "):]
        if "SignedDistanceField.java" in f:                                                      # J10
            s = re.sub(r"public\s+(int|short)\s+(\w+)\(", r"virtual \1 \2(", s)
        sdfs.append(s)

    consts = read("Constants.java")
    cl = []
    for name in CONSTANTS:
        m = re.search(r"public static final (\w+) %s = ([^;]+);" % name, consts)
        assert m, name
        cl.append("  static inline %s %s = %s;" % ({"byte": "jbyte"}.get(m.group(1), m.group(1)), name, m.group(2)))
    cl.append("  static inline int SUB_OCTREE_SIZE = 512;  // OctreeThread.java: constructInnerOctree(512, ...)  (J9)")
    cl.append("  static inline int SDF_MAX_LOD = 13;       // Octree.java: useSDFBrush(..., 0, 13, changeBounds)   (J9)")

    parts = ["// GENERATED by oracle/build_ref_java.py from /root/reference/src/engine/*.java -- do not commit.",
             "struct Constants {\n" + "\n".join(cl) + "\n};",
             "struct Octree; struct SignedDistanceField; struct OctreeThread;",
             close_classes(rewrite(util)),
             "\n".join(close_classes(rewrite(s)) for s in sdfs),
             "struct Octree {\n" + close_classes(oct_cpp) + "\n};"]
    # OctreeThread needs the complete Octree and Octree::buildChunk the complete OctreeThread: buildChunk is defined last.
    parts.append(close_classes(rewrite(thread)))
    parts.append("inline " + rewrite(splice).replace("void buildChunk(Chunk* chunk", "void Octree::buildChunk(Octree::Chunk* chunk", 1))
    return "\n\n".join(parts) + "\n"


def generate_voxeliser() -> str:
    """src/shaders/chunkgen-heightmap.comp through rules R1-R3 of oracle/build_ref.py (layout/uniform lines become variables,
    the rest becomes the body of `struct Invocation`, float literals get `f`), plus: the literal `2048` (= WORLD_SIZE / 4 of the
    shipped 8192^3 world) becomes `ref_height_scale`, default 2048 (J9)."""
    from build_ref import rewrite_common, split_items
    path = os.path.join(os.path.dirname(REF), "shaders", "chunkgen-heightmap.comp")
    src = strip_comments(open(path, encoding="utf-8", errors="replace").read())
    hoisted, body = [], []
    for it in split_items(src):
        if it.startswith("#version"):
            continue
        m = re.match(r"layout\s*\(([^)]*)\)\s*(.*)$", it, flags=re.S)
        rest = m.group(2) if m else it
        if m and re.match(r"in\s*;", rest):
            continue
        if rest.startswith("uniform"):
            hoisted.append(re.sub(r"^uniform\s+", "", rest))
            continue
        assert not m, "unhandled layout item: " + it[:60]
        body.append(rewrite_common(it))
    text = "\n".join(body)
    assert len(re.findall(r"\* 2048\b", text)) == 1
    text = re.sub(r"\* 2048\b", "* ref_height_scale", text)
    return ("// GENERATED by oracle/build_ref_java.py from %s -- do not commit (oracle/_ref/ is git-ignored)\n"
            "namespace glslv { namespace ref_chunkgen {\nint ref_height_scale = 2048;\n%s\nstruct Invocation : invocation_base {\n%s\n};\n} }\n"
            % (path, "\n".join(hoisted), text))


def build(force: bool = False) -> str | None:
    """Returns the library path, or None when neither /root/reference nor a prebuilt library is there."""
    srcs = [os.path.join(HERE, f) for f in ("java_shim.h", "glsl_voxel_shim.h", "ref_java_harness.cpp", "build_ref_java.py", "build_ref.py")]
    java = [os.path.join(REF, f) for f in ("Octree.java", "OctreeThread.java", "Util.java", "Constants.java", "sdf/Sphere.java",
                                           "sdf/Box.java", "sdf/SignedDistanceField.java")]
    if not os.path.isfile(java[0]):
        return LIB if os.path.exists(LIB) else None
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs + java):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    with open(GEN, "w") as f:
        f.write(generate())
    with open(os.path.join(OUT, "ref_chunkgen_gen.inc"), "w") as f:
        f.write(generate_voxeliser())
    subprocess.check_call(["g++", *CXXFLAGS, "-I", HERE, "-I", OUT, "-o", LIB, os.path.join(HERE, "ref_java_harness.cpp"), "-lm"])
    return LIB


if __name__ == "__main__":
    if "--print" in sys.argv:
        sys.stdout.write(generate())
    else:
        p = build(force="--force" in sys.argv)
        print(p or "no /root/reference and no prebuilt oracle/_ref/libsvo_ref_java.so")
