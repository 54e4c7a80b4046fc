#!/usr/bin/env python3
"""Builds oracle/_ref/libsvo_ref.so: THE REFERENCE'S OWN SHADERS, compiled for the CPU.  TEST INFRASTRUCTURE.

The text of /root/reference/src/shaders/svotrace.comp and svobeam.comp is read where it lies (nothing is copied into
the repository: the generated C++ goes to oracle/_ref/, which is git-ignored), put through the MECHANICAL rewrite
below -- syntax only -- and compiled by g++ against oracle/glsl_shim.h (the GLSL vocabulary as C++) and
oracle/ref_harness.cpp (the dispatch loop + extern "C" entry points).  Control flow, expression order, constants,
record decode, traversal and shading are therefore the reference's, parsed by a C++ compiler; only what GLSL leaves
to the GL driver (rounding of + - * /, min/max of NaN, sin/cos/acos/exp, undefined reads) is fixed by the shim, to
the same contract as oracle/oracle_math.h.

The rewrite, in full (each step is a regular expression over the comment-stripped text):
  R1  `#version`, `layout(local_size...) in;` dropped; `layout(...) uniform T x;` / `uniform T x;` -> a namespace-scope
      variable `T x;`; `layout(std430...) [readonly] buffer B { uint[] x; };` -> `ssbo_uint x;` (bounds-checked: U1).
  R2  everything else (functions, structs, globals, #defines) becomes the body of `struct Invocation`, so that the
      shader's mutable globals (`debugColor`, `octstack`, `stack_ptr`) are per-invocation state as in GLSL.
  R3  floating literals without suffix get `f` (GLSL literals are float; C++ would compute in double).
  R4  parameter qualifiers: `out T x` / `inout T x` -> `T& x`; `in T x` -> `T x`.
  R5  `float rand = rand(...)` (svotrace.comp:486): the local that shadows the function is renamed `rand_local`.
  R6  `getByte(nodePointer++)`: the k-th occurrence inside one function becomes `getByte(nodePointer + k)` -- the
      left-to-right operand evaluation the shader relies on, which C++ does not promise for `|`.
  R7  single scalar declarations without initialiser (`float beamDist;`, struct fields) get `= {}` (U2: zero).
  R8  instrumentation that changes no value: `intersectOctree` is renamed `intersectOctree_glsl` and wrapped by a
      logger appended to the struct (records hit flag, res.pointer / t / value / iter per cast); `iter++;` also
      copies the counter to `probe_iter` (the shader's own store of it is commented out, svotrace.comp:728).
  R9  three constants become variables WHOSE DEFAULTS ARE THE SHIPPED VALUES, so that the oracle's generalisations
      can be pinned too: `#define MAX_DEPTH 13` -> ref_max_depth, the cone cut `maxDepth = 11` -> ref_cone_depth,
      the mode-0 loop bound `i < 2` -> ref_casts.
  R10 `octstack[MAX_SCALE + 1]` is declared with 33 entries: a POP after `pos` left [1,2) reads index up to 31 (the
      value is never used: the loop exits as a miss); GL robust access makes that read harmless, C++ would not.

Usage: python oracle/build_ref.py [--force]     (needs /root/reference; the GPU box uses the prebuilt .so)
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SHADERS = "/root/reference/src/shaders"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libsvo_ref.so")
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
            "-fno-unsafe-math-optimizations", "-fno-strict-aliasing", "-w"]


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def split_items(src: str):
    """Top-level items of a GLSL translation unit: preprocessor lines, declarations (end at `;` at depth 0),
    function definitions (end at the `}` that closes them)."""
    items, i, n = [], 0, len(src)
    while i < n:
        while i < n and src[i].isspace():
            i += 1
        if i >= n:
            break
        if src[i] == "#":
            j = src.find("\n", i)
            j = n if j < 0 else j
            items.append(src[i:j].rstrip())
            i = j
            continue
        j, depth, first_brace = i, 0, -1
        while j < n:
            c = src[j]
            if c == "{":
                if depth == 0 and first_brace < 0:
                    first_brace = j
                depth += 1
            elif c == "}":
                depth -= 1
                if depth == 0:
                    head = src[i:first_brace]
                    is_decl = re.match(r"\s*(struct|layout|uniform|const)\b", head) or "=" in head
                    if not is_decl:      # function definition ends here
                        j += 1
                        break
            elif c == ";" and depth == 0:
                j += 1
                break
            j += 1
        items.append(src[i:j].strip())
        i = j
    return items


FLOAT_LIT = re.compile(r"(?<![\w.])(\d+\.\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def rewrite_common(text: str) -> str:
    text = FLOAT_LIT.sub(lambda m: m.group(1) + "f", text)                                   # R3
    text = re.sub(r"(?<=[(,])(\s*)(?:out|inout)\s+(\w+)\s+(\w+)", r"\1\2& \3", text)          # R4
    text = re.sub(r"(?<=[(,])(\s*)in\s+(\w+)\s+(\w+)", r"\1\2 \3", text)
    return text


def rewrite_item(item: str) -> str:
    k = [0]

    def number(_m):                                                                          # R6
        k[0] += 1
        return "getByte(nodePointer + %du)" % (k[0] - 1)

    item = re.sub(r"getByte\(\s*nodePointer\+\+\s*\)", number, item)
    item = re.sub(r"\brand\b(?!\s*\()", "rand_local", item)                                  # R5
    item = re.sub(r"(?m)^(\s*)(float|int|uint|bool)\s+(\w+)\s*;", r"\1\2 \3 = {};", item)    # R7
    return item


def generate(shader: str, namespace: str) -> str:
    src = strip_comments(open(os.path.join(REF_SHADERS, shader), encoding="utf-8", errors="replace").read())
    hoisted, body = [], []
    for it in split_items(src):
        if it.startswith("#version"):
            continue                                                                          # R1
        if it.startswith("#"):
            if re.match(r"#define\s+MAX_DEPTH\b", it):
                it = "#define MAX_DEPTH ref_max_depth"                                        # R9
            body.append(rewrite_common(it))
            continue
        m = re.match(r"layout\s*\(([^)]*)\)\s*(.*)$", it, flags=re.S)
        rest = m.group(2) if m else it
        if m and re.match(r"in\s*;", rest):
            continue                                                                          # R1 local_size
        if re.match(r"(readonly\s+)?buffer\b", rest):
            name = re.search(r"uint\s*\[\s*\]\s*(\w+)\s*;", rest).group(1)
            hoisted.append(("ssbo_uint %s;" if rest.startswith("readonly") else "ssbo_uint_rw %s;") % name)
            continue
        if rest.startswith("uniform"):
            hoisted.append(rewrite_common(re.sub(r"^uniform\s+", "", rest)))
            continue
        assert not m, "unhandled layout item: " + it[:60]
        body.append(rewrite_item(rewrite_common(it)))
    text = "\n".join(body)
    # R8
    n_def = len(re.findall(r"\bbool\s+intersectOctree\s*\(", text))
    assert n_def == 1, n_def
    text = re.sub(r"\bbool\s+intersectOctree\s*\(", "bool intersectOctree_glsl(", text)
    assert len(re.findall(r"\biter\+\+;", text)) == 1
    text = re.sub(r"\biter\+\+;", "iter++; probe_iter = iter;", text)
    # R9
    if namespace == "ref_svotrace":
        assert len(re.findall(r"maxDepth = 11;", text)) == 1
        text = text.replace("maxDepth = 11;", "maxDepth = ref_cone_depth;")
        assert len(re.findall(r"i\s*<\s*2\s*;", text)) == 1
        text = re.sub(r"i\s*<\s*2\s*;", "i < ref_casts;", text)
    # R10
    assert len(re.findall(r"octstack\[MAX_SCALE \+ 1\]", text)) == 1
    text = text.replace("octstack[MAX_SCALE + 1]", "octstack[33]")
    wrapper = """
  /* R8: logger around the renamed shader function (build_ref.py; not reference code) */
  bool intersectOctree(vec3 origin, vec3 dir, vec3 invdir, castResult& res, int maxDepth, bool coneTrace){
    probe_iter = 0;
    bool hit = intersectOctree_glsl(origin, dir, invdir, res, maxDepth, coneTrace);
    if(n_casts < REF_MAX_LOG){
      cast_log[n_casts].hit = hit; cast_log[n_casts].loop_iter = probe_iter; cast_log[n_casts].pointer = res.pointer;
      cast_log[n_casts].t = res.t; cast_log[n_casts].value = res.value; cast_log[n_casts].iter = res.iter;
    }
    n_casts++;
    return hit;
  }
"""
    return ("// GENERATED by oracle/build_ref.py from %s/%s -- do not commit (oracle/_ref/ is git-ignored)\n"
            "namespace glsl { namespace %s {\nint ref_max_depth = 13, ref_cone_depth = 11, ref_casts = 2;\n%s\n"
            "struct Invocation : ref_invocation_base {\n%s\n%s};\n} }\n"
            % (REF_SHADERS, shader, namespace, "\n".join(hoisted), text, wrapper))


def build(force: bool = False) -> str | None:
    """Returns the library path, or None when neither /root/reference nor a prebuilt library is there."""
    srcs = [os.path.join(HERE, f) for f in ("glsl_shim.h", "ref_harness.cpp", "oracle_math.h", "build_ref.py")]
    have_ref = os.path.isfile(os.path.join(REF_SHADERS, "svotrace.comp"))
    if not have_ref:
        return LIB if os.path.exists(LIB) else None
    shaders = [os.path.join(REF_SHADERS, s) for s in ("svotrace.comp", "svobeam.comp")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs + shaders):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    for shader, ns in (("svotrace.comp", "ref_svotrace"), ("svobeam.comp", "ref_svobeam")):
        with open(os.path.join(OUT, ns + "_gen.inc"), "w") as f:
            f.write(generate(shader, ns))
    subprocess.check_call(["g++", *CXXFLAGS, "-I", HERE, "-I", OUT, "-o", LIB, os.path.join(HERE, "ref_harness.cpp"),
                           "-lpthread", "-lm"])
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p or "no /root/reference and no prebuilt oracle/_ref/libsvo_ref.so")
