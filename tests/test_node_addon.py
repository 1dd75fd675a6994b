"""The Node.js N-API addon (host_node/rt2015_napi.c) EXECUTED without Node.js: the addon is
linked with an in-process stand-in for the N-API subset it uses (host_node/test/napi_mock.c) and driven the way a
JavaScript host would drive it (host_node/test/addon_test.c).  CPU: every header entry point is exported, the
generated wrappers are up to date, loaders / struct probes / argument errors / the loud no-device failure work.
GPU: an Assignment-1 frame and a small Assignment-10 render made THROUGH the addon equal the golden fixture resp.
the same inputs replayed through the ctypes binding, bit for bit."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import golden_io as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE = os.path.join(ROOT, "host_node")


@pytest.fixture(scope="module")
def addon_test(rt, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("addon") / "addon_test")
    libdir = os.path.dirname(rt.lib.LIB_PATH)
    cmd = ["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + NODE, "-o", exe,
           os.path.join(NODE, "test", "addon_test.c"), os.path.join(NODE, "test", "napi_mock.c"), os.path.join(NODE, "rt2015_napi.c"),
           "-L" + libdir, "-lrt2015", "-Wl,-rpath," + libdir, "-lm"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    return exe


def test_generated_wrappers_are_current():
    p = subprocess.run([sys.executable, os.path.join(NODE, "gen_napi.py"), "--check"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout


def test_addon_exports_the_whole_abi_and_runs_on_the_cpu(addon_test):
    p = subprocess.run([addon_test, "cpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    out = p.stdout
    exported = sorted(re.findall(r"^export (\w+)$", out, re.M))
    with open(os.path.join(ROOT, "include", "rt2015.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = sorted(set(re.findall(r"^\s*(?:int|unsigned|void\s*\*|void|const\s+char\s*\*)\s*rt_([A-Za-z0-9_]+)\s*\(", text, re.M)))
    internal = {"mesh_data_free", "mol_data_free"}   # called by parse_mesh_json / parse_pdb after copying into TypedArrays
    assert exported == [d for d in declared if d not in internal]
    assert "struct_size Ray10 48 Poi10 64 Poi8 48 Ray6 48 Nope 0" in out
    # serial numbers 1,2,(TER 3),4 -> atoms.length 4 with 3 records (quirk Q13); C, O, N -> 3 elements
    assert "parse_pdb size 4 records 3 elements 3 atomData.length 12 last 0.000 -2.000 1.000" in out
    assert "parse_mesh_json triangles 2 materials 1 positions.length 18 p[3] 1.0 p[16] 1.0 material 0.500 0.125" in out
    assert "struct_size(1 arg): TypeError: rt2015: wrong number of arguments" in out
    assert "ctx_create: ok" in out or "ctx_create threw: Error: no CUDA device (there is no CPU fallback)" in out


def _f(path, dtype):
    return np.fromfile(path, dtype=dtype)


@pytest.mark.gpu
def test_addon_renders_like_the_ctypes_binding(rt, addon_test, tmp_path):
    p = subprocess.run([addon_test, "gpu", str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0 and "gpu ok" in p.stdout, p.stdout
    d = lambda n, t: _f(os.path.join(str(tmp_path), n), t)   # noqa: E731
    # Assignment 1 against the golden fixture (the reference's own kernel)
    assert np.array_equal(d("a01_pixels.bin", np.uint8).reshape(64, 64, 4), G.load("a01")["pixels_64x64"])
    # Assignment 10: the same inputs through ctypes
    L = rt.lib
    cols, rows, rpp = 48, 32, 4
    total = cols * rows * rpp
    # every host array is bound to a name first: a temporary would be freed (and its memory reused) before the C call reads it
    f32 = {k: d("in_%s.bin" % k, np.float32) for k in ("bound", "materials", "sbound", "tbound", "l_shadow", "l_scene", "l_light", "cam")}
    f64 = {k: d("in_%s.bin" % k, np.float64) for k in ("xyzr", "pos9", "nor9", "smin", "smax", "tmin", "tmax")}
    with L.Context(0) as ctx:
        gs, gt = L.Grid(), L.Grid()
        sid, tid = np.array([0], np.uint32), np.array([1, 1], np.uint32)
        d3 = {k: (C.c_double * 3)(*f64[k]) for k in ("smin", "smax", "tmin", "tmax")}
        ctx.check(L.dll.rt_grid_build_spheres(ctx.h, L.hptr(f64["xyzr"]), L.hptr(sid), 1, d3["smin"], d3["smax"], 1, C.byref(gs)))
        ctx.check(L.dll.rt_grid_build_triangles(ctx.h, L.hptr(f64["pos9"]), L.hptr(f64["nor9"]), L.hptr(tid), 2, d3["tmin"], d3["tmax"], 1, None,
                                                C.byref(gt)))
        scene = C.c_void_p()
        ctx.check(L.dll.rt_scene_create(ctx.h, C.byref(scene)))
        ctx.check(L.dll.rt_scene_set_bounds(scene, L.hptr(f32["bound"])))
        ctx.check(L.dll.rt_scene_set_materials(scene, L.hptr(f32["materials"]), 2))
        ctx.check(L.dll.rt_scene_add_set(scene, C.byref(gs), L.hptr(f32["sbound"]), 0, 0))
        ctx.check(L.dll.rt_scene_add_set(scene, C.byref(gt), L.hptr(f32["tbound"]), 0, 0))
        ctx.check(L.dll.rt_scene_add_light(scene, L.hptr(f32["l_shadow"]), L.hptr(f32["l_scene"]), L.hptr(f32["l_light"])))
        opts = L.RenderOpts()
        opts.cols, opts.rows, opts.rays_per_pixel, opts.depth = cols, rows, rpp, 5
        opts.focal_length, opts.lens_rad = 5.0, np.float32(0.05)
        render = C.c_void_p()
        ctx.check(L.dll.rt_render_create(ctx.h, scene, C.byref(opts), C.byref(render)))
        seeds = d("in_seeds.bin", np.int32)
        ctx.check(L.dll.rt_render_set_seeds(render, L.hptr(seeds), total, 0))
        cam = f32["cam"]
        pix = np.zeros(cols * rows * 4, np.uint8)
        for _ in range(2):
            ctx.check(L.dll.rt_render_execute(render, L.hptr(cam), L.hptr(pix)))
        acc = np.zeros(cols * rows * 4, np.float32)
        ctx.check(L.dll.rt_render_read_accum(render, L.hptr(acc)))
        after = np.zeros(total, np.int32)
        ctx.check(L.dll.rt_render_read_seeds(render, L.hptr(after), total))
        ctx.check(L.dll.rt_render_destroy(render))
        ctx.check(L.dll.rt_scene_destroy(scene))
        L.dll.rt_grid_release(ctx.h, C.byref(gt))
        L.dll.rt_grid_release(ctx.h, C.byref(gs))
    assert pix.reshape(-1, 4)[:, :3].any(), "the frame is not empty"
    assert np.array_equal(d("a10_pixels.bin", np.uint8), pix)
    assert np.array_equal(d("a10_accum.bin", np.uint32), acc.view(np.uint32))
    assert np.array_equal(d("a10_seeds.bin", np.int32), after)
    assert not np.array_equal(after, seeds), "the pass drew random numbers"
