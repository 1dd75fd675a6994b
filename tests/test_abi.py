"""CPU tests (no GPU, no compute calls): the C-ABI library loads and exports exactly the entry
points include/rt2015.h declares, the ctypes binding covers all of them, and the product fails
loudly -- never falls back -- when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rt2015.h")
DECL = re.compile(r"^\s*(?:int|unsigned|void\s*\*|void|const\s+char\s*\*)\s*(rt_[A-Za-z0-9_]+)\s*\(", re.M)


def declared():
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(DECL.findall(text)))


def test_header_declares_the_kernel_surface():
    names = declared()
    # one launcher per kernel of the reference, named after it
    for k in ("initAcu", "initTrace", "bouncePaths", "lightRender", "initShadowTrace", "sphereTrace", "triangleTrace", "meshTrace",
              "sphereShadowTrace", "triangleShadowTrace", "sceneRender", "copyToPixel"):
        assert "rt_a10_" + k in names
    for k in ("rt_ctx_create", "rt_buffer_create", "rt_buffer_write", "rt_buffer_read", "rt_struct_size", "rt_grid_build_spheres",
              "rt_grid_build_triangles", "rt_render_create", "rt_render_execute", "rt_finish"):
        assert k in names


def test_library_exports_every_declared_symbol(rt):
    dll = C.CDLL(rt.lib.LIB_PATH)
    missing = [n for n in declared() if not hasattr(dll, n)]
    assert not missing, "declared in include/rt2015.h but not exported: %s" % missing


def test_binding_covers_the_header(rt):
    assert sorted(rt.lib.EXPORTS) == declared()


def test_struct_sizes(rt):
    """sizeofRay / sizeofPoi probes (A10/code.cl:440-446) -- host-only function, no device needed."""
    f = rt.lib.dll.rt_struct_size
    assert f(b"Ray", 10) == 48 and f(b"Poi", 10) == 64
    assert f(b"Poi", 8) == 48 and f(b"Poi", 9) == 48 and f(b"Ray", 7) == 48 and f(b"Ray", 3) == 48
    assert f(b"Poi", 7) == 0 and f(b"Nope", 10) == 0


def test_no_cpu_fallback(rt):
    """Without a CUDA device rt_ctx_create must fail with RT_ERR_NO_DEVICE and the Python layer
    must raise; with one it must succeed.  There is no third outcome."""
    h = C.c_void_p()
    rc = rt.lib.dll.rt_ctx_create(0, C.byref(h))
    assert rc in (0, -4)
    if rc == 0:
        assert rt.lib.dll.rt_ctx_destroy(h) == 0
    else:
        assert not h.value
        with pytest.raises(rt.lib.RtError):
            rt.lib.Context(0)
        with pytest.raises(rt.lib.RtError):
            rt.Renderer({"spheres": [], "triangles": [], "meshes": [], "lights": []}, 8, 8)


def test_null_arguments_are_rejected(rt):
    dll = rt.lib.dll
    assert dll.rt_ctx_create(0, None) == -1
    assert dll.rt_finish(None) == -1
    assert dll.rt_buffer_create(None, 16, None) == -1
    assert dll.rt_render_execute(None, None, None) == -1
    assert dll.rt_last_error_string(None) == b"null context"


def test_comm_id_needs_no_gpu():
    """rt_comm_unique_id binds NCCL at run time (dlopen) and needs neither a context nor a device: it either hands out a 128-byte id or
    fails with a status -- never a crash -- and rejects a null pointer before touching NCCL.  In a process of its own: binding the
    system's libnccl.so.2 here would shadow the newer one PyTorch bundles for every later `import torch` of the test session."""
    import subprocess
    import sys
    code = ("import importlib, ctypes\n"
            "rt = importlib.import_module('2015-raytracing_b200')\n"
            "f = rt.lib.dll.rt_comm_unique_id\n"
            "f.restype = ctypes.c_int\n"
            "assert f(None) == -1\n"
            "buf = (ctypes.c_ubyte * 128)()\n"
            "rc = f(buf)\n"
            "assert rc <= 0 and (rc != 0 or any(bytes(buf)))\n"
            "print('comm id ok', rc)\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "comm id ok" in r.stdout, r.stdout


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import or load it."""
    pkg = os.path.join(ROOT, "2015-raytracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn), encoding="utf-8") as f:
                    text = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), fn
                assert "libref" not in text and "librt_oracle" not in text, fn


def test_napi_addon_compiles():
    """The Node N-API addon of INTEGRATION.md type-checks against include/rt2015.h and the
    hand-declared N-API subset (Node.js itself is absent from the image)."""
    import subprocess
    src = os.path.join(ROOT, "host_node", "rt2015_napi.c")
    p = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "host_node"), src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
