"""C ABI checks that need no GPU: the library loads, exports every symbol include/svo_b200.h declares, struct
layouts agree between header, product and oracle, and the product refuses to run without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "svo_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(svo):
    names = _declared()
    assert len(names) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", svo._lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (svo_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, "declared in svo_b200.h but not exported: %s" % missing
    assert sorted(svo._lib.SYMBOLS) == names, "the ctypes table and the header disagree"
    lib = svo._lib.lib()
    assert lib.svo_abi_version() == 1


def test_frame_layout_matches_header_and_oracle(svo, oracle):
    assert C.sizeof(svo.Frame) == 15 * 4 + 8 * 4 == 92
    assert [f[0] for f in svo.Frame._fields_] == [f[0] for f in oracle.Frame._fields_]
    for name, _ in svo.Frame._fields_:
        assert getattr(svo.Frame, name).offset == getattr(oracle.Frame, name).offset
    assert svo.RAY_DTYPE.itemsize == 24 and svo.HIT_DTYPE.itemsize == 16
    assert oracle.RAY_DTYPE == svo.RAY_DTYPE and oracle.HIT_DTYPE == svo.HIT_DTYPE


def _have_gpu(svo):
    n = C.c_int()
    return svo._lib.lib().svo_device_count(C.byref(n)) == 0 and n.value > 0


def test_no_cpu_fallback(svo):
    """Without a CUDA device the product path fails loudly (SVO_ERR_NO_DEVICE); it never computes on the CPU."""
    if _have_gpu(svo):
        pytest.skip("a GPU is present")
    with pytest.raises(svo.SvoError) as e:
        svo.SvoContext(64, 64)
    assert e.value.code == svo._lib.ERR_NO_DEVICE
    with pytest.raises(svo.SvoError):
        svo.Renderer(64, 64)


def test_argument_validation_without_device(svo):
    lib = svo._lib.lib()
    h = C.c_void_p()
    assert lib.svo_create(C.byref(h), 0, 0, 10) == svo._lib.ERR_INVALID
    assert lib.svo_create(None, 0, 10, 10) == svo._lib.ERR_INVALID
    assert b"" != lib.svo_last_error(None)
    assert lib.svo_upload(None, None, 0) == svo._lib.ERR_INVALID
    assert lib.svo_render(None, None) == svo._lib.ERR_INVALID
    lib.svo_destroy(None)  # no-op
    need = C.c_uint64()
    hm = np.zeros((8, 8), np.uint16)
    mm = np.ones((8, 8), np.uint8)
    # world generation is host code: sizes and fills without a device
    assert lib.svo_build_terrain(hm.ctypes.data_as(C.c_void_p), mm.ctypes.data_as(C.c_void_p), 8, 8, None, 0, C.byref(need), 1) == 0
    assert need.value > 7
    assert lib.svo_build_terrain(hm.ctypes.data_as(C.c_void_p), mm.ctypes.data_as(C.c_void_p), 6, 8, None, 0, C.byref(need), 1) == svo._lib.ERR_INVALID


def _build_c_client(tmp_path):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "engine_loop")
    libdir = os.path.join(root, "svo_raytracer_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c_client", "engine_loop.c"), "-o", exe, "-L", libdir, "-lsvo_b200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_c_client_compiles_links_and_refuses_to_run_without_a_device(tmp_path):
    """include/svo_b200.h bound by a C compiler (the Java binding of INTEGRATION.md cannot be compiled here): the engine's
    frame-loop call sequence (tests/c_client/engine_loop.c) compiles as C11 with -Werror and links every entry point it
    needs; without a CUDA device it exits 77 -- svo_create fails loudly, there is no CPU path."""
    import subprocess
    import svo_raytracer_b200 as svo
    import ctypes as C
    exe = _build_c_client(tmp_path)
    n = C.c_int(0)
    svo._lib.lib().svo_device_count(C.byref(n))
    rc = subprocess.run([exe], capture_output=True, text=True)
    assert rc.returncode == (0 if n.value > 0 else 77), (rc.returncode, rc.stdout, rc.stderr)


@pytest.mark.gpu
def test_c_client_runs_the_engine_loop(tmp_path):
    import subprocess
    rc = subprocess.run([_build_c_client(tmp_path)], capture_output=True, text=True)
    assert rc.returncode == 0 and "engine loop ok" in rc.stdout, (rc.returncode, rc.stdout, rc.stderr)


def _build_cpp_client(tmp_path):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "engine_main")
    libdir = os.path.join(root, "svo_raytracer_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp_client", "engine_main.cpp"), "-o", exe, "-L", libdir, "-lsvo_b200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_cpp_renderer_mirror_compiles_and_fails_loudly_without_a_device(tmp_path):
    """include/svo_renderer.hpp -- Renderer.java's interface (addShader / addSSBO / updateSSBO / dispatchCompute /
    printGLErrors, Renderer.java:43-165) restated in C++ over the C ABI -- compiles with -Werror together with the engine's
    frame loop written against it (tests/cpp_client/engine_main.cpp).  Without a CUDA device createImages queues
    SVO_ERR_NO_DEVICE and the process stays alive (exit 77): no exception, no abort, no CPU path."""
    import subprocess
    import svo_raytracer_b200 as svo
    import ctypes as C
    exe = _build_cpp_client(tmp_path)
    n = C.c_int(0)
    svo._lib.lib().svo_device_count(C.byref(n))
    rc = subprocess.run([exe], capture_output=True, text=True)
    assert rc.returncode == (0 if n.value > 0 else 77), (rc.returncode, rc.stdout, rc.stderr)


@pytest.mark.gpu
def test_cpp_renderer_mirror_draws_the_frames_of_the_bare_abi(tmp_path):
    """The engine loop through svo::Renderer (modes 2, 0, 3, 1, both beam pre-passes, an edit pushed as two updateSSBO
    ranges) equals the same frames drawn through the bare C ABI byte for byte; its error queue behaves like glGetError."""
    import subprocess
    rc = subprocess.run([_build_cpp_client(tmp_path)], capture_output=True, text=True)
    assert rc.returncode == 0 and "engine main ok" in rc.stdout, (rc.returncode, rc.stdout, rc.stderr)
    assert "Update SSBO error: Invalid parameters." in rc.stdout
