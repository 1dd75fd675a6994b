"""GPU parity tests of the Assignment-10 path (grid build, every kernel, the frame) against
the oracle on the same seeded inputs, all through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import util
from oracle import host as OH
from oracle import refcl as OR

pytestmark = pytest.mark.gpu

COLS, ROWS, RPP = 96, 64, 4


@pytest.fixture(scope="module")
def scenes(tmp_path_factory):
    return util.make_scene_pair(tmp_path_factory.mktemp("a10"), COLS, ROWS, mesh_uv=(32, 16), mesh_nslabs=10)


def _grid_arrays(ctx, rt, g):
    d = rt.host.DeviceGrid(ctx, g, np.zeros(8, np.float32))
    out = {"box": d.box_size(), "prim": d.prim(), "matid": d.matid()}
    if g.kind == 1 and g.normal:
        out["normal"] = d.normal()
    return out


def test_grid_build_bit_exact(rt, gpu_ctx, scenes):
    """Cell lists (box_size prefix sums + cell-ordered primitive buffers) must match the
    reference's JS split*Data bit for bit -- BASELINE.md gate 1."""
    o_scene, p_scene = scenes
    prep = OR.prepare_a10(o_scene, 1)
    # scene triangles / spheres at n_slabs = 1 (A10 default) and at 3 and 5
    for n in (1, 3, 5):
        data, mat, box = OH.splitSphereData(o_scene, n)
        g = rt.splitSphereData(gpu_ctx, p_scene, n)
        got = _grid_arrays(gpu_ctx, rt, g)
        assert np.array_equal(got["box"], box)
        assert np.array_equal(got["prim"].view(np.uint32), OH.to_f32(data).view(np.uint32))
        assert np.array_equal(got["matid"], mat)
        rt.lib.dll.rt_grid_release(gpu_ctx.h, C.byref(g))
        pos, nor, mat, box = OH.splitTriangleData(o_scene, n)
        g = rt.splitTriangleData(gpu_ctx, p_scene, n)
        got = _grid_arrays(gpu_ctx, rt, g)
        assert np.array_equal(got["box"], box)
        assert np.array_equal(got["prim"].view(np.uint32), OH.to_f32(pos).view(np.uint32))
        assert np.array_equal(got["normal"].view(np.uint32), OH.to_f32(nor).view(np.uint32))
        assert np.array_equal(got["matid"], mat)
        rt.lib.dll.rt_grid_release(gpu_ctx.h, C.byref(g))
    # the <mesh>: per-mesh nslabs, transform applied after the split in float64
    om, pm = o_scene["meshes"][0], p_scene["meshes"][0]
    g = pm.upload(gpu_ctx)
    got = _grid_arrays(gpu_ctx, rt, g)
    assert np.array_equal(got["box"], np.asarray(om.boxSizeData, np.uint32))
    assert np.array_equal(got["prim"].view(np.uint32), OH.to_f32(om.posData).view(np.uint32))
    assert np.array_equal(got["normal"].view(np.uint32), OH.to_f32(om.normalData).view(np.uint32))
    assert np.array_equal(rt.bounds2AABB(pm.bounds).view(np.uint32), OH.bounds2AABB(om.bounds).view(np.uint32))
    assert got["box"][-1] > om.ntriangles  # triangles really are duplicated across cells


def _upload_prep(ctx, prep):
    dev = {"materials": ctx.upload(prep["materials"]), "sets": []}
    for s in prep["sets"]:
        d = {"kind": s["kind"], "box": ctx.upload(s["box"]), "aabb": s["aabb"], "n": s["n"]}
        if s["kind"] == "sphere":
            d["data"] = ctx.upload(s["data"])
            d["matid"] = ctx.upload(s["matid"])
        else:
            d["pos"] = ctx.upload(s["pos"])
            d["normal"] = ctx.upload(s["normal"])
            d["matid"] = ctx.upload(s["matid"]) if s["kind"] == "triangle" else s["matid"]
        dev["sets"].append(d)
    return dev


def _hp(a):
    return a.ctypes.data


def test_every_kernel_matches_oracle(rt, gpu_ctx, oracle_lib, scenes):
    """Drive the reference's executeRender schedule kernel by kernel through the C ABI on the
    GPU and through the oracle on the CPU, on the same uploaded buffers, and compare the
    device state after EVERY launch."""
    o_scene, _ = scenes
    ctx, dll = gpu_ctx, rt.lib.dll
    prep = OR.prepare_a10(o_scene, 1)
    dev = _upload_prep(ctx, prep)
    total = COLS * ROWS * RPP
    seeds0 = OR.make_seeds(total, 7)
    st = OR.A10State(total, seeds0)
    oracle_lib.a10_initAcu(st.acu, total)
    d_rays, d_pois, d_shadow = ctx.alloc(48 * total), ctx.alloc(64 * total), ctx.alloc(48 * total)
    d_acu, d_seeds = ctx.alloc(16 * total), ctx.upload(seeds0)
    ctx.call("rt_buffer_fill", d_rays, 0, 48 * total)
    ctx.call("rt_buffer_fill", d_pois, 0, 64 * total)
    ctx.call("rt_buffer_fill", d_shadow, 0, 48 * total)
    ctx.call("rt_a10_initAcu", d_acu, total)
    cam = o_scene["camera"].toFloat32Array()
    focal = float(np.float32(o_scene["focal_length"]))
    lens = float(np.float32(o_scene["lens_diameter"] / 2.0))
    report = []

    def compare(tag):
        rays = ctx.download(d_rays, OR.RAY, total)
        pois = ctx.download(d_pois, OR.POI10, total)
        shadow = ctx.download(d_shadow, OR.RAY, total)
        acu = ctx.download(d_acu, np.float32, total * 4).reshape(-1, 4)
        seeds = ctx.download(d_seeds, np.int32, total)
        assert np.array_equal(seeds, st.seeds), tag + ": seed state differs (RNG draw count/order)"
        live = st.rays["mint"] != st.rays["maxt"]
        assert np.array_equal(live, rays["mint"] != rays["maxt"]), tag + ": live-ray sets differ"
        worst = 0
        for name, a, b in (("ray.o", rays["o"][live, :3], st.rays["o"][live, :3]), ("ray.d", rays["d"][live, :3], st.rays["d"][live, :3]),
                           ("ray.mint", rays["mint"], st.rays["mint"]), ("ray.maxt", rays["maxt"], st.rays["maxt"])):
            worst = max(worst, int(util.ulp_diff(a, b).max(initial=0)))
        hit = st.pois["matId"] >= 0
        assert np.array_equal(pois["matId"], st.pois["matId"]), tag + ": matId differs"
        for f in ("p", "normal", "atte"):
            worst = max(worst, int(util.ulp_diff(pois[f][hit, :3], st.pois[f][hit, :3]).max(initial=0)))
        slive = st.shadow["mint"] != st.shadow["maxt"]
        assert np.array_equal(slive, shadow["mint"] != shadow["maxt"]), tag + ": lit/blocked sets differ"
        worst = max(worst, int(util.ulp_diff(shadow["d"][slive, :3], st.shadow["d"][slive, :3]).max(initial=0)))
        worst = max(worst, int(util.ulp_diff(acu, st.acu).max(initial=0)))
        report.append((tag, worst))
        assert worst == 0, "%s: max ulp distance %d (expected bit-exact)" % (tag, worst)

    def closest(tag):
        for s, d in zip(prep["sets"], dev["sets"]):
            if s["kind"] == "sphere":
                oracle_lib.a10_sphereTrace(total, st.pois, st.rays, s["data"], s["matid"], s["box"], s["aabb"], s["n"])
                ctx.call("rt_a10_sphereTrace", total, d_pois, d_rays, d["data"], d["matid"], d["box"], _hp(s["aabb"]), s["n"])
            elif s["kind"] == "triangle":
                oracle_lib.a10_triangleTrace(total, st.pois, st.rays, s["pos"], s["normal"], s["matid"], s["box"], s["aabb"], s["n"])
                ctx.call("rt_a10_triangleTrace", total, d_pois, d_rays, d["pos"], d["normal"], d["matid"], d["box"], _hp(s["aabb"]), s["n"])
            else:
                oracle_lib.a10_meshTrace(total, st.pois, st.rays, s["pos"], s["normal"], s["box"], s["matid"], s["aabb"], s["n"])
                ctx.call("rt_a10_meshTrace", total, d_pois, d_rays, d["pos"], d["normal"], d["box"], s["matid"], _hp(s["aabb"]), s["n"])
            compare(tag + ":" + s["kind"])

    def shade(tag):
        for li, L in enumerate(prep["lights"]):
            oracle_lib.a10_initShadowTrace(st.shadow, st.pois, total, L["shadow"], st.seeds)
            ctx.call("rt_a10_initShadowTrace", d_shadow, d_pois, total, _hp(L["shadow"]), d_seeds)
            compare("%s:initShadow%d" % (tag, li))
            for s, d in zip(prep["sets"], dev["sets"]):
                if s["kind"] == "sphere":
                    oracle_lib.a10_sphereShadowTrace(total, st.shadow, s["data"], s["box"], s["aabb"], s["n"])
                    ctx.call("rt_a10_sphereShadowTrace", total, d_shadow, d["data"], d["box"], _hp(s["aabb"]), s["n"])
                else:
                    oracle_lib.a10_triangleShadowTrace(total, st.shadow, s["pos"], s["box"], s["aabb"], s["n"])
                    ctx.call("rt_a10_triangleShadowTrace", total, d_shadow, d["pos"], d["box"], _hp(s["aabb"]), s["n"])
            oracle_lib.a10_sceneRender(st.acu, st.pois, st.shadow, prep["materials"], L["scene"], total)
            ctx.call("rt_a10_sceneRender", d_acu, d_pois, d_shadow, dev["materials"], _hp(L["scene"]), total)
            compare("%s:sceneRender%d" % (tag, li))

    oracle_lib.a10_initTrace(st.seeds, st.rays, st.pois, prep["aabb"], cam, focal, lens, RPP, COLS, ROWS, 0)
    ctx.call("rt_a10_initTrace", d_seeds, d_rays, d_pois, _hp(prep["aabb"]), _hp(cam), focal, lens, RPP)
    compare("initTrace")
    closest("primary")
    for li, L in enumerate(prep["lights"]):
        oracle_lib.a10_lightRender(st.pois, st.rays, st.acu, L["light"], total)
        ctx.call("rt_a10_lightRender", d_pois, d_rays, d_acu, _hp(L["light"]), total)
        compare("lightRender%d" % li)
    shade("primary")
    for j in range(5):
        oracle_lib.a10_bouncePaths(st.pois, st.rays, st.seeds, total)
        ctx.call("rt_a10_bouncePaths", d_pois, d_rays, d_seeds, total)
        compare("bounce%d" % j)
        closest("bounce%d" % j)
        shade("bounce%d" % j)
    pix_o = np.zeros((COLS * ROWS, 4), np.uint8)
    oracle_lib.a10_copyToPixel(pix_o, st.acu, 1.0 / RPP, COLS * ROWS, RPP)
    d_pix = ctx.alloc(4 * COLS * ROWS)
    ctx.call("rt_a10_copyToPixel", d_pix, d_acu, 1.0 / RPP, COLS * ROWS, RPP)
    pix = ctx.download(d_pix, np.uint8, 4 * COLS * ROWS).reshape(-1, 4)
    assert np.array_equal(pix, pix_o)
    assert pix[:, :3].max() > 30, "image is not trivially black"


@pytest.mark.parametrize("mode", [1, 0, 2])
def test_render_frame_matches_oracle(rt, oracle_lib, scenes, mode):
    """rt_render_execute (mode 1 = reference schedule, mode 0 = wavefront stages + pair-list queue
    walkers, mode 2 = megakernel) vs the
    oracle's executeRender: per-pixel float accumulation within 1e-3 (BASELINE.md gate 4; in
    practice bit-exact), seed buffer equal as integers, two progressive passes."""
    o_scene, p_scene = scenes
    total = COLS * ROWS * RPP
    seeds0 = OR.make_seeds(total, 11)
    prep = OR.prepare_a10(o_scene, 1)
    st = OR.A10State(total, seeds0)
    oracle_lib.a10_initAcu(st.acu, total)
    cam = o_scene["camera"].toFloat32Array()
    r = rt.Renderer(p_scene, COLS, ROWS, RPP, mode=mode, tile_slots=COLS * 8 * RPP)
    r.preRender(seeds0)
    try:
        for p in range(2):
            pix_o = OR.a10_execute_render(oracle_lib, st, prep, cam, COLS, ROWS, RPP, o_scene["focal_length"], o_scene["lens_diameter"])
            pix = r.executeRender()
            acc = r.accum()
            acc_o = st.acu.reshape(COLS * ROWS, RPP, 4)
            ref = np.zeros((COLS * ROWS, 4), np.float32)
            for k in range(RPP):
                ref += acc_o[:, k]
            scale = 1.0 / (RPP * (p + 1))
            assert np.abs(acc[:, :3] - ref[:, :3]).max() * scale <= 1e-3
            assert np.array_equal(r.seeds(), st.seeds)
            assert np.array_equal(acc.view(np.uint32), ref.view(np.uint32)), "expected bit-exact accumulation"
            assert np.array_equal(pix.reshape(-1, 4), pix_o.reshape(-1, 4))
            s = r.stats()
            assert s["launches"] > 0
        assert r.stats()["closest_rays"] + r.stats()["any_rays"] > 0
        assert st.n_closest + st.n_any > 0
    finally:
        r.postRender()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_render_frame_multi_cell_xml_sets(rt, oracle_lib, scenes, mode):
    """Global n_slabs = 3 (the reference's commented-out UI input, A10/index.html:28): the XML sphere and triangle
    sets become multi-cell grids too and are served by the queue walkers (sphere and triangle instantiations)."""
    o_scene, p_scene = scenes
    total = COLS * ROWS * RPP
    seeds0 = OR.make_seeds(total, 23)
    prep = OR.prepare_a10(o_scene, 3)
    st = OR.A10State(total, seeds0)
    oracle_lib.a10_initAcu(st.acu, total)
    cam = o_scene["camera"].toFloat32Array()
    OR.a10_execute_render(oracle_lib, st, prep, cam, COLS, ROWS, RPP, o_scene["focal_length"], o_scene["lens_diameter"])
    r = rt.Renderer(p_scene, COLS, ROWS, RPP, n_slabs=3, mode=mode)
    r.preRender(seeds0)
    try:
        r.executeRender()
        ref = np.zeros((COLS * ROWS, 4), np.float32)
        for k in range(RPP):
            ref += st.acu.reshape(COLS * ROWS, RPP, 4)[:, k]
        assert np.array_equal(r.seeds(), st.seeds)
        assert np.array_equal(r.accum().view(np.uint32), ref.view(np.uint32))
    finally:
        r.postRender()


def test_full_size_frame_rows_match_oracle(rt, oracle_lib, tmp_path):
    """BASELINE config 5 geometry at full size -- 1920x1080, the synthetic 1 000 000-triangle <mesh> at
    nslabs 128 -- rendered whole on the GPU; the oracle renders a band of pixel rows of the same frame
    (every kernel is per-slot independent, so a row tile is exact) and must agree bit for bit on the
    accumulation image and on the seed buffer of those rows.  Also checks the two size-independent
    invariances of the pipeline: the result does not depend on the wavefront tile size, nor on how
    the slots of a pixel are split between renderers (the multi-GPU partition)."""
    import synth
    cols, rows, rpp = 1920, 1080, 4
    mesh_json = synth.synth_mesh(1000, 500, seed=2015)
    path = synth.write_scene(tmp_path, n_lights=2, with_sphere=True, with_mesh=True, mesh_nslabs=128)
    o_scene = OH.loadScene(path, cols, rows, mesh_loader=lambda _f: OH.parseMeshJSON(mesh_json))
    p_scene = rt.loadScene(path, cols, rows, mesh_loader=lambda _f: rt.parseMeshJSON(mesh_json))
    total = cols * rows * rpp
    seeds0 = OR.make_seeds(total, 2015)
    r = rt.Renderer(p_scene, cols, rows, rpp)
    r.preRender(seeds0)
    try:
        # the GPU grid build of the 1 M-triangle mesh against the oracle's cell lists
        om, g = o_scene["meshes"][0], p_scene["meshes"][0].grid
        d = rt.host.DeviceGrid(r.ctx, g, np.zeros(8, np.float32))
        assert np.array_equal(d.box_size(), np.asarray(om.boxSizeData, np.uint32))
        assert np.array_equal(d.prim().view(np.uint32), OH.to_f32(om.posData).view(np.uint32))
        r.executeRender(readback=False)
        acc = r.accum().reshape(rows, cols, 4)
        seeds1 = r.seeds().reshape(rows, cols * rpp)
        stats = r.stats()
    finally:
        r.postRender()
    assert stats["closest_rays"] > 6 * 0.5 * cols * rows * rpp
    prep = OR.prepare_a10(o_scene, 1)
    cam = o_scene["camera"].toFloat32Array()
    for row0, nrows in ((0, 2), (537, 6), (1078, 2)):
        band = seeds0.reshape(rows, cols * rpp)[row0:row0 + nrows].reshape(-1)
        st = OR.A10State(cols * nrows * rpp, band)
        oracle_lib.a10_initAcu(st.acu, st.total)
        OR.a10_execute_render(oracle_lib, st, prep, cam, cols, rows, rpp, o_scene["focal_length"], o_scene["lens_diameter"], row0=row0, nrows=nrows)
        ref = np.zeros((cols * nrows, 4), np.float32)
        for k in range(rpp):
            ref += st.acu.reshape(cols * nrows, rpp, 4)[:, k]
        got = acc[row0:row0 + nrows].reshape(-1, 4)
        assert np.abs(got[:, :3] - ref[:, :3]).max() / rpp <= 1e-3
        assert np.array_equal(seeds1[row0:row0 + nrows].reshape(-1), st.seeds), "RNG streams differ in rows %d.." % row0
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "expected bit-exact rows %d.." % row0
    # tile-size invariance and slot-split invariance (bit-exact seeds; accumulation: same per-slot values, the
    # per-pixel sum of a split render is the sum of the partial sums)
    p2 = rt.loadScene(path, cols, rows, mesh_loader=lambda _f: rt.parseMeshJSON(mesh_json))
    ra = rt.Renderer(p2, cols, rows, rpp, tile_slots=1 << 20)
    ra.preRender(seeds0)
    try:
        ra.executeRender(readback=False)
        assert np.array_equal(ra.accum().view(np.uint32).reshape(rows, cols, 4), acc.view(np.uint32))
        assert np.array_equal(ra.seeds().reshape(rows, cols * rpp), seeds1)
        ctx = ra.ctx
        parts, accs = [], []
        for rank in range(2):
            b, c = rt.multi.slot_range(rank, 2, rpp)
            p3 = rt.loadScene(path, cols, rows, mesh_loader=lambda _f: rt.parseMeshJSON(mesh_json))
            rb = rt.Renderer(p3, cols, rows, rpp, slots=(b, c), ctx=ctx)
            rb.preRender(seeds0)
            try:
                rb.executeRender(readback=False)
                parts.append((b, c, rb.seeds()))
                accs.append(rb.accum())
            finally:
                rb.postRender()
        assert np.array_equal(rt.multi.merge_seeds(parts, cols * rows, rpp), seeds1.reshape(-1))
        both = accs[0] + accs[1]
        assert np.abs(both - acc.reshape(-1, 4)).max() <= 1e-5 * max(1.0, float(np.abs(acc).max()))
    finally:
        ra.postRender()


@pytest.mark.parametrize("mode", [0, 1])
def test_rays_per_pixel_one_serial_outcome(rt, oracle_lib, scenes, mode):
    """rays_per_pixel == 1: the reference draws the lens sample from seeds[get_global_id(0)] = seeds[col] in a 2-D
    launch, so all rows of a column race on one seed (quirk Q7).  The only defined outcome is the serial row-major
    one; we implement it (column `col` hands draws 2*row, 2*row+1 of its stream to row `row`) and the oracle runs
    the same kernel text single-threaded in that order."""
    o_scene, p_scene = scenes
    total = COLS * ROWS
    seeds0 = OR.make_seeds(total, 31)
    prep = OR.prepare_a10(o_scene, 1)
    st = OR.A10State(total, seeds0)
    oracle_lib.a10_initAcu(st.acu, total)
    cam = o_scene["camera"].toFloat32Array()
    r = rt.Renderer(p_scene, COLS, ROWS, 1, mode=mode)
    r.preRender(seeds0)
    try:
        for _ in range(2):
            pix_o = OR.a10_execute_render(oracle_lib, st, prep, cam, COLS, ROWS, 1, o_scene["focal_length"], o_scene["lens_diameter"], serial_init=True)
            pix = r.executeRender()
            assert np.array_equal(r.seeds(), st.seeds)
            assert np.array_equal(r.accum().view(np.uint32), st.acu.view(np.uint32))
            assert np.array_equal(pix.reshape(-1, 4), pix_o.reshape(-1, 4))
    finally:
        r.postRender()


def test_error_paths_and_empty_launches(rt, gpu_ctx, scenes):
    """Argument checking of the C ABI on a live context: bad arguments give RT_ERR_INVALID / RT_ERR_STATE with a
    message, zero-sized launches are no-ops, a render refuses to run before its seeds are set."""
    ctx, dll = gpu_ctx, rt.lib.dll
    d = ctx.alloc(64)
    assert dll.rt_a10_initAcu(ctx.h, d, 0) == 0
    assert dll.rt_a10_bouncePaths(ctx.h, d, d, d, 0) == 0
    assert dll.rt_a10_copyToPixel(ctx.h, d, d, 1.0, 0, 4) == 0
    assert dll.rt_a10_initAcu(ctx.h, None, 4) == -1
    assert dll.rt_a10_sphereTrace(ctx.h, 4, d, d, d, d, d, None, 1) == -1
    assert dll.rt_a10_sphereTrace(ctx.h, 4, d, d, d, d, d, np.zeros(8, np.float32).ctypes.data, 0) == -1
    assert dll.rt_buffer_write(ctx.h, d, 0, 16, None) == -1
    _, p_scene = scenes
    with pytest.raises(rt.lib.RtError, match="perfect square"):
        r = rt.Renderer(p_scene, COLS, ROWS, 3, ctx=ctx)
        r.preRender(None)
    r = rt.Renderer(p_scene, 8, 8, 4, ctx=ctx)   # canvas differs from the scene camera's cols/rows
    r.preRender(np.arange(1, 8 * 8 * 4 + 1, dtype=np.int32))
    try:
        with pytest.raises(rt.lib.RtError, match="camera cols/rows"):
            r.executeRender()
    finally:
        r.postRender()
    r = rt.Renderer(p_scene, COLS, ROWS, 4, ctx=ctx)
    r.preRender(None)
    try:
        with pytest.raises(rt.lib.RtError, match="seeds not set"):
            r.executeRender()
        with pytest.raises(rt.lib.RtError, match="count must be"):
            r.setSeeds(np.ones(5, np.int32))
        r.setSeeds(np.arange(1, COLS * ROWS * 4 + 1, dtype=np.int32))
        img = r.executeRender()
        assert img.shape == (ROWS, COLS, 4) and img[..., 3].min() == 255
    finally:
        r.postRender()
    with pytest.raises(rt.lib.RtError, match="slot range"):
        r2 = rt.Renderer(p_scene, COLS, ROWS, 4, slots=(3, 2), ctx=ctx)
        try:
            r2.preRender(None)
        finally:
            r2.postRender()


def test_checkpoint_resume_is_bit_exact(rt, scenes, tmp_path):
    """Progressive display path: 3 passes in one go == 1 pass, export (acu, seeds, passes), import into a fresh
    render, 2 more passes -- accumulation, seeds and the displayed image identical; the image goes out as a PNG."""
    _, p_scene = scenes
    total = COLS * ROWS * RPP
    seeds0 = OR.make_seeds(total, 41)
    a = rt.Renderer(p_scene, COLS, ROWS, RPP)
    a.preRender(seeds0)
    try:
        for _ in range(3):
            img_a = a.executeRender()
        acc_a, seeds_a = a.accum(), a.seeds()
        ctx = a.ctx
        b = rt.Renderer(p_scene, COLS, ROWS, RPP, ctx=ctx)
        b.preRender(seeds0)
        try:
            b.executeRender()
            state = b.export_state()
            assert state["passes"] == 2
        finally:
            b.postRender()
        c = rt.Renderer(p_scene, COLS, ROWS, RPP, ctx=ctx)
        c.preRender(None)
        try:
            c.import_state(state)
            c.executeRender()
            img_c = c.executeRender()
            assert np.array_equal(c.seeds(), seeds_a)
            assert np.array_equal(c.accum().view(np.uint32), acc_a.view(np.uint32))
            assert np.array_equal(img_c, img_a)
        finally:
            c.postRender()
    finally:
        a.postRender()
    png = tmp_path / "frame.png"
    rt.write_png(str(png), img_a)
    from PIL import Image
    back = np.asarray(Image.open(str(png)).convert("RGBA"))
    assert np.array_equal(back, img_a)


def test_sincos_equals_separate_sin_and_cos(tmp_path):
    """The concentric map's cos/sin pair comes from one sincos() call on the GPU (RT_SINCOS in rt_device.cuh); the arithmetic
    contract says double-precision cos and sin rounded to fp32.  Exhaustive over every fp32 angle with |x| <= 4."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / "sincos_check")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sincos_check.cu")
    b = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-o", exe, src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert b.returncode == 0, b.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout


def test_fast_triangle_test_equals_reference_order(tmp_path):
    """interTriangleFast rejects on the sign of the barycentric numerators BEFORE the division; the reference (A10/code.cl:262-275)
    divides first, and `beta < 0` does not fire when the quotient underflows to -0.  The early exit is guarded (|numerator| >= 2^-126,
    div <= 2^22); this checks 2e8 cases incl. subnormal numerators against div up to 2^120 / +inf, and that the corner was exercised."""
    import os
    import re
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "tri_fast_check")
    b = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
                        "-I", os.path.join(here, "..", "2015-raytracing_b200", "csrc"), "-o", exe, os.path.join(here, "tri_fast_check.cu")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert b.returncode == 0, b.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout
    m = re.search(r"accepted (\d+) underflow_corner (\d+)", r.stdout)
    assert m and int(m.group(1)) > 1000 and int(m.group(2)) > 1000, r.stdout


@pytest.mark.parametrize("mode,rpp", [(0, 4), (0, 1), (2, 4), (1, 4)])
def test_non_blocking_seed_upload_is_equivalent(rt, scenes, mode, rpp):
    """rt_render_write_local_seeds_async (the reference's non-blocking enqueueWriteBuffer) overlaps the upload with the head of the
    next pass; whichever kernel reads the seeds first must see them: every path (wavefront, rpp == 1, megakernel, reference
    schedule) gives the results of the blocking upload, pass after pass, also when the seeds are replaced between passes."""
    import torch
    _, p_scene = scenes
    total = COLS * ROWS * rpp
    seeds = [OR.make_seeds(total, 100 + k) for k in range(3)]
    pinned = [torch.from_numpy(s.copy()).pin_memory() for s in seeds]
    out = []
    for non_blocking in (False, True):
        r = rt.Renderer(p_scene, COLS, ROWS, rpp, mode=mode)
        r.preRender(seeds[0])
        res = []
        try:
            for k in range(3):
                src = pinned[k].numpy() if non_blocking else seeds[k]
                r.writeLocalSeeds(src, non_blocking=non_blocking)
                pix = r.executeRender()
                res.append((pix.copy(), r.accum(), r.seeds()))
        finally:
            r.postRender()
        out.append(res)
    for (p0, a0, s0), (p1, a1, s1) in zip(*out):
        assert np.array_equal(s0, s1) and np.array_equal(a0.view(np.uint32), a1.view(np.uint32)) and np.array_equal(p0, p1)


def test_nccl_reduce_from_the_library(rt, scenes):
    """rt_comm_* / rt_render_reduce: the accumulation image is reduced by an ncclReduce the C library issues on the context's
    stream (libnccl bound at run time).  One GPU here, so a communicator of one rank: the call path (dlopen, unique id, init,
    reduce on the render stream, destroy) runs for real and must leave the image untouched; the N > 1 sum is covered by the
    2/4/8-GPU bench runs and, for the split / merge logic, by tests/test_multi_gloo.py."""
    _, p_scene = scenes
    r = rt.Renderer(p_scene, COLS, ROWS, RPP)
    r.preRender(OR.make_seeds(COLS * ROWS * RPP, 5))
    comm = None
    try:
        r.executeRender()
        before = r.accum()
        comm = rt.multi.Comm(r.ctx, 0, 1, exchange=lambda ident: ident)
        comm.reduce(r, 0)
        r.ctx.finish()
        assert np.array_equal(r.accum().view(np.uint32), before.view(np.uint32))
        with pytest.raises(rt.lib.RtError):
            comm.reduce(r, 3)   # root outside the communicator
    finally:
        if comm is not None:
            comm.close()
        r.postRender()


@pytest.mark.parametrize("mode", [0, 2])
def test_caller_built_grids_without_occupancy_bits(rt, scenes, monkeypatch, mode):
    """The drop-in ABI accepts grids the CALLER built (any box_size / prim buffers; `occupancy` is an internal extra of the
    library's own builder).  Multi-cell grids that arrive without occupancy bits get them derived from the cell table in
    rt_scene_add_set; the frame must equal the one rendered from the library-built grids (global n_slabs 3: sphere and
    triangle sets are multi-cell too, next to the mesh)."""
    _, p_scene = scenes
    seeds0 = OR.make_seeds(COLS * ROWS * RPP, 77)

    def frame():
        r = rt.Renderer(p_scene, COLS, ROWS, RPP, n_slabs=3, mode=mode)
        r.preRender(seeds0)
        try:
            r.executeRender()
            return r.accum(), r.seeds()
        finally:
            r.postRender()

    want_acc, want_seeds = frame()
    real = rt.lib.dll.rt_scene_add_set
    stripped = []

    def add_set_without_occupancy(hs, grid_ref, bound, is_mesh, matid):
        g = grid_ref._obj
        assert g.n_slabs > 1 and g.occupancy
        bare = rt.lib.Grid.from_buffer_copy(bytes(g))
        bare.occupancy = None
        stripped.append(bare)
        return real(hs, C.byref(bare), bound, is_mesh, matid)

    monkeypatch.setattr(rt.lib.dll, "rt_scene_add_set", add_set_without_occupancy, raising=False)
    got_acc, got_seeds = frame()
    assert len(stripped) == 3
    assert np.array_equal(got_seeds, want_seeds)
    assert np.array_equal(got_acc.view(np.uint32), want_acc.view(np.uint32))


def test_empty_slot_range_is_rejected(rt, scenes, gpu_ctx):
    """multi.slot_range hands (begin, 0) to ranks beyond rays_per_pixel; such a rank must not silently render EVERY slot
    (slot_count 0 used to mean "all"): the Python layer and the C ABI both refuse."""
    _, p_scene = scenes
    assert rt.multi.slot_range(5, 8, 4) == (4, 0)
    with pytest.raises(ValueError, match="empty slot range"):
        rt.Renderer(p_scene, COLS, ROWS, 4, slots=(4, 0), ctx=gpu_ctx)
    r = rt.Renderer(p_scene, COLS, ROWS, 4, ctx=gpu_ctx)
    r.slots = (2, 0)   # past the Python check, straight to rt_render_create
    with pytest.raises(rt.lib.RtError, match="empty slot range"):
        try:
            r.preRender(None)
        finally:
            r.postRender()
