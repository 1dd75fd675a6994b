"""GPU parity gates of BASELINE.md section 4 that need the reference's INSTRUMENTED kernels (oracle/_ref/libref_instr.so,
hooks at non-arithmetic places of the unmodified code.cl, oracle/cl2cpp.py) or the configs' own sampling density:

* hit-primitive-id gate: the winner `champ_i` of the reference's per-cell loop (A10/code.cl:882-897, A07/code.cl:402-424,
  541-567), per work-item, against rt_set_walk_stats' hit ids -- >= 99.99 % first, then equality;
* work counters C (cells visited) / T (primitive tests) / H (hits) the roofline numerator stands on:
  per work-item from the launchers, and the fused path's `profile_pass()` totals, against the oracle's counters;
* BASELINE config 5 at its own sampling (256 slots per pixel = 16 x 16 lens grid) and A09 at its default 100 rays per
  pixel (10 x 10), against the oracle on the same inputs.
"""
import ctypes as C

import numpy as np
import pytest

import golden_io as G
import util
from oracle import host as OH
from oracle import refcl as OR

pytestmark = pytest.mark.gpu

NONE = 0xFFFFFFFF


@pytest.fixture(scope="module")
def instr_lib():
    if not OR.have_reference():
        pytest.skip("oracle/_ref (the reference's own kernels, instrumented build) is not built here")
    lib = OR.load_reference(instrumented=True)
    assert lib.instrumented
    return lib


class _Sinks:
    """Per-work-item statistics buffers on both sides."""

    def __init__(self, ctx, lib, n):
        self.ctx, self.lib, self.n = ctx, lib, n
        self.o_hit = np.zeros(n, np.uint32)
        self.o_cells = np.zeros(n, np.uint64)
        self.o_tests = np.zeros(n, np.uint64)
        self.d_hit, self.d_cells, self.d_tests = ctx.alloc(4 * n), ctx.alloc(4 * n), ctx.alloc(4 * n)
        lib.set_stats(self.o_hit, self.o_cells, self.o_tests)
        ctx.call("rt_set_walk_stats", self.d_hit, self.d_cells, self.d_tests)

    def gpu(self):
        return (self.ctx.download(self.d_hit, np.uint32, self.n), self.ctx.download(self.d_cells, np.uint32, self.n),
                self.ctx.download(self.d_tests, np.uint32, self.n))

    def check(self, tag):
        hit, cells, tests = self.gpu()
        agree = float((hit == self.o_hit).mean())
        assert agree >= 0.9999, "%s: hit primitive ids agree on %.4f %% of the work-items" % (tag, 100 * agree)
        assert np.array_equal(hit, self.o_hit), "%s: hit primitive ids (champ_i) differ" % tag
        assert np.array_equal(cells.astype(np.uint64), self.o_cells), "%s: cells visited differ" % tag
        assert np.array_equal(tests.astype(np.uint64), self.o_tests), "%s: primitive tests differ" % tag
        return hit, cells, tests

    def close(self):
        self.lib.set_stats(None, None, None)
        self.ctx.call("rt_set_walk_stats", None, None, None)
        for p in (self.d_hit, self.d_cells, self.d_tests):
            self.ctx.free(p)


@pytest.mark.parametrize("name", G.names("tri_") + G.names("mol_"))
def test_hit_primitive_ids_a07(rt, gpu_ctx, instr_lib, tmp_path, name):
    """Config 3 (and the molecule twin): per-pixel winner of A07's meshTrace / molTrace on the reference's own meshes and
    molecules at every grid resolution of the fixture, plus the per-pixel cell and test counts."""
    fx = G.load(name)
    P = fx["params"]
    cols, rows = P["cols"], P["rows"]
    if name.startswith("tri_"):
        m = G.meshes_of(fx)[0]
        p = tmp_path / "m.json"
        p.write_text(G.mesh_json_text(m["positions"], m["normals"], m["materialIndices"], m["materials"]))
        o_data, p_data, kw = OH.parseMeshJSON(str(p)), rt.parseMeshJSON(str(p)), "meshData"
    else:
        text = G.pdb_text(fx["serial"], fx["elem"], fx["xyz"])
        o_data, p_data, kw = OH.parsePDB(text), rt.parsePDB(text), "molData"
    s = _Sinks(gpu_ctx, instr_lib, cols * rows)
    try:
        some_hit = False
        for n in P["slabs"]:
            _, rays_o, _ = OR.a07_render(instr_lib, cols, rows, n, **{kw: o_data})
            _, maxt = rt.assignments.a07_compute(gpu_ctx, cols, rows, n, **{kw: p_data})
            hit, cells, _ = s.check("%s n=%d" % (name, n))
            assert np.array_equal(maxt.view(np.uint32), rays_o["maxt"].view(np.uint32))
            some_hit = some_hit or bool((hit != NONE).any())
            assert cells.sum() > 0
        assert some_hit
    finally:
        s.close()


COLS, ROWS, RPP = 96, 64, 4


def test_hit_ids_and_work_counters_a10(rt, gpu_ctx, instr_lib, tmp_path):
    """One whole Assignment-10 pass (primary segment + 5 bounces, two lights) kernel by kernel on both sides with the
    statistics sinks on: every grid-walk launch must report the reference's champ_i, cell count and test count per ray
    slot.  The per-set totals of those launches are then what the fused path's instrumented pass (`profile_pass`, the
    numerator of bench.py's roofline) must report for the same scene and seeds."""
    o_scene, p_scene = util.make_scene_pair(tmp_path, COLS, ROWS, mesh_uv=(32, 16), mesh_nslabs=10)
    ctx = gpu_ctx
    prep = OR.prepare_a10(o_scene, 1)
    dev = {"materials": ctx.upload(prep["materials"]), "sets": []}
    for st_ in prep["sets"]:
        d = {"box": ctx.upload(st_["box"])}
        if st_["kind"] == "sphere":
            d["data"], d["matid"] = ctx.upload(st_["data"]), ctx.upload(st_["matid"])
        else:
            d["pos"], d["normal"] = ctx.upload(st_["pos"]), ctx.upload(st_["normal"])
            d["matid"] = ctx.upload(st_["matid"]) if st_["kind"] == "triangle" else st_["matid"]
        dev["sets"].append(d)
    total = COLS * ROWS * RPP
    seeds0 = OR.make_seeds(total, 31)
    st = OR.A10State(total, seeds0)
    lib = instr_lib
    lib.a10_initAcu(st.acu, total)
    d_rays, d_pois, d_shadow = ctx.alloc(48 * total), ctx.alloc(64 * total), ctx.alloc(48 * total)
    d_acu, d_seeds = ctx.alloc(16 * total), ctx.upload(seeds0)
    for p, b in ((d_rays, 48), (d_pois, 64), (d_shadow, 48)):
        ctx.call("rt_buffer_fill", p, 0, b * total)
    ctx.call("rt_a10_initAcu", d_acu, total)
    cam = o_scene["camera"].toFloat32Array()
    focal, lens = float(np.float32(o_scene["focal_length"])), float(np.float32(o_scene["lens_diameter"] / 2.0))
    hp = lambda a: a.ctypes.data   # noqa: E731
    n_sets = len(prep["sets"])
    want = np.zeros((n_sets, 16), np.uint64)   # rows of rt_render_read_profile_sets, from the ORACLE's counters
    s = _Sinks(ctx, lib, total)

    def tally(k, base, alive, hit, cells, tests, kind):
        want[k, base + 0] += int(alive)
        want[k, base + 1] += int((cells > 0).sum())
        want[k, base + 2] += int(cells.sum())
        want[k, base + (3 if kind == "sphere" else 4)] += int(tests.sum())
        nh = int((hit != NONE).sum())
        if base == 0:
            want[k, 5 if kind == "sphere" else (6 if kind == "triangle" else 7)] += nh
        else:
            want[k, 13] += nh

    def closest(tag):
        for k, (q, d) in enumerate(zip(prep["sets"], dev["sets"])):
            alive = np.count_nonzero(st.rays["mint"] != st.rays["maxt"])
            if q["kind"] == "sphere":
                lib.a10_sphereTrace(total, st.pois, st.rays, q["data"], q["matid"], q["box"], q["aabb"], q["n"])
                ctx.call("rt_a10_sphereTrace", total, d_pois, d_rays, d["data"], d["matid"], d["box"], hp(q["aabb"]), q["n"])
            elif q["kind"] == "triangle":
                lib.a10_triangleTrace(total, st.pois, st.rays, q["pos"], q["normal"], q["matid"], q["box"], q["aabb"], q["n"])
                ctx.call("rt_a10_triangleTrace", total, d_pois, d_rays, d["pos"], d["normal"], d["matid"], d["box"], hp(q["aabb"]), q["n"])
            else:
                lib.a10_meshTrace(total, st.pois, st.rays, q["pos"], q["normal"], q["box"], q["matid"], q["aabb"], q["n"])
                ctx.call("rt_a10_meshTrace", total, d_pois, d_rays, d["pos"], d["normal"], d["box"], q["matid"], hp(q["aabb"]), q["n"])
            hit, cells, tests = s.check("%s closest %s" % (tag, q["kind"]))
            tally(k, 0, alive, hit, cells, tests, q["kind"])

    def shade(tag):
        for li, L in enumerate(prep["lights"]):
            lib.a10_initShadowTrace(st.shadow, st.pois, total, L["shadow"], st.seeds)
            ctx.call("rt_a10_initShadowTrace", d_shadow, d_pois, total, hp(L["shadow"]), d_seeds)
            for k, (q, d) in enumerate(zip(prep["sets"], dev["sets"])):
                alive = np.count_nonzero(st.shadow["mint"] != st.shadow["maxt"])
                if q["kind"] == "sphere":
                    lib.a10_sphereShadowTrace(total, st.shadow, q["data"], q["box"], q["aabb"], q["n"])
                    ctx.call("rt_a10_sphereShadowTrace", total, d_shadow, d["data"], d["box"], hp(q["aabb"]), q["n"])
                else:
                    lib.a10_triangleShadowTrace(total, st.shadow, q["pos"], q["box"], q["aabb"], q["n"])
                    ctx.call("rt_a10_triangleShadowTrace", total, d_shadow, d["pos"], d["box"], hp(q["aabb"]), q["n"])
                hit, cells, tests = s.check("%s light %d any-hit %s" % (tag, li, q["kind"]))
                tally(k, 8, alive, hit, cells, tests, q["kind"])
            lib.a10_sceneRender(st.acu, st.pois, st.shadow, prep["materials"], L["scene"], total)
            ctx.call("rt_a10_sceneRender", d_acu, d_pois, d_shadow, dev["materials"], hp(L["scene"]), total)

    try:
        lib.a10_initTrace(st.seeds, st.rays, st.pois, prep["aabb"], cam, focal, lens, RPP, COLS, ROWS, 0)
        ctx.call("rt_a10_initTrace", d_seeds, d_rays, d_pois, hp(prep["aabb"]), hp(cam), focal, lens, RPP)
        closest("primary")
        for L in prep["lights"]:
            lib.a10_lightRender(st.pois, st.rays, st.acu, L["light"], total)
            ctx.call("rt_a10_lightRender", d_pois, d_rays, d_acu, hp(L["light"]), total)
        shade("primary")
        for j in range(5):
            lib.a10_bouncePaths(st.pois, st.rays, st.seeds, total)
            ctx.call("rt_a10_bouncePaths", d_pois, d_rays, d_seeds, total)
            closest("bounce %d" % j)
            shade("bounce %d" % j)
        assert np.array_equal(ctx.download(d_seeds, np.int32, total), st.seeds)
        assert np.array_equal(ctx.download(d_acu, np.float32, 4 * total).view(np.uint32), st.acu.reshape(-1).view(np.uint32))
    finally:
        s.close()
        for p in (d_rays, d_pois, d_shadow, d_acu, d_seeds):
            ctx.free(p)
    want[0, 14] = total
    assert want[:, 2].sum() > 0 and want[:, 4].sum() > 0 and want[:, 5:8].sum() > 0 and want[:, 13].sum() > 0
    # the fused path's instrumented pass on the same scene and seeds: the roofline's C / T / H
    r = rt.Renderer(p_scene, COLS, ROWS, RPP, ctx=ctx)
    r.preRender(seeds0)
    try:
        got = r.profile_pass()
    finally:
        r.postRender()
    assert np.array_equal(got[:n_sets, :15], want[:, :15]), "work counters of the instrumented pass differ from the oracle's:\n%s\n%s" % (got[:n_sets], want)
    assert not got[n_sets:].any()


@pytest.fixture(scope="module")
def config5(rt, tmp_path_factory):
    """BASELINE config 5 geometry: 1920x1080, synthetic 1 000 000-triangle <mesh> at nslabs 128, two disk lights."""
    import synth
    tmp = tmp_path_factory.mktemp("cfg5")
    cols, rows = 1920, 1080
    mesh_json = synth.synth_mesh(1000, 500, seed=2015)
    path = synth.write_scene(tmp, n_lights=2, with_sphere=True, with_mesh=True, mesh_nslabs=128)
    o_scene = OH.loadScene(path, cols, rows, mesh_loader=lambda _f: OH.parseMeshJSON(mesh_json))
    p_scene = rt.loadScene(path, cols, rows, mesh_loader=lambda _f: rt.parseMeshJSON(mesh_json))
    return cols, rows, o_scene, p_scene


def test_config5_at_256_slots_per_pixel(rt, oracle_lib, config5):
    """The bench workload itself -- 1920x1080, rays_per_pixel 256 (16 x 16 stratified lens grid, A10/code.cl:482-508), depth 5,
    1 M-triangle mesh -- rendered whole on the GPU (530 841 600 slots, two wavefront tiles); the oracle renders two pixel rows
    of the same frame (2 x 1920 x 256 = 983 040 slots) and must agree bit for bit on accumulation and seed state."""
    cols, rows, o_scene, p_scene = config5
    rpp = 256
    rng = np.random.Generator(np.random.PCG64(2015))
    check_rows = (269, 811)
    # the seed array of the whole frame is 2.1 GB: fill it in row bands, remember the checked rows
    seeds = np.empty(cols * rows * rpp, np.int32)
    band = cols * rpp
    for row in range(rows):
        seeds[row * band:(row + 1) * band] = rng.integers(1, 2 ** 31, size=band, dtype=np.int64).astype(np.int32)
    r = rt.Renderer(p_scene, cols, rows, rpp)
    r.preRender(seeds)
    try:
        r.executeRender(readback=False)
        acc = r.accum().reshape(rows, cols, 4)
        stats = r.stats()
        seeds_after = r.seeds()
        got_seeds = {row: seeds_after[row * band:(row + 1) * band].copy() for row in check_rows}
        del seeds_after
    finally:
        r.postRender()
    assert stats["closest_rays"] + stats["any_rays"] > 8_000_000_000
    prep = OR.prepare_a10(o_scene, 1)
    cam = o_scene["camera"].toFloat32Array()
    for row in check_rows:
        st = OR.A10State(band, seeds[row * band:(row + 1) * band])
        oracle_lib.a10_initAcu(st.acu, st.total)
        OR.a10_execute_render(oracle_lib, st, prep, cam, cols, rows, rpp, o_scene["focal_length"], o_scene["lens_diameter"], row0=row, nrows=1)
        ref = np.zeros((cols, 4), np.float32)
        for k in range(rpp):
            ref += st.acu.reshape(cols, rpp, 4)[:, k]
        got = acc[row]
        assert np.abs(got[:, :3] - ref[:, :3]).max() / rpp <= 1e-3, "row %d: accumulation" % row
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "row %d: expected bit-exact accumulation" % row
        assert np.array_equal(got_seeds[row], st.seeds), "row %d: RNG streams differ" % row
        assert (ref[:, 3] > 0).any()


@pytest.mark.parametrize("name", [n for n in G.names("a09_")][:3])
def test_a09_at_default_100_rays_per_pixel(rt, gpu_ctx, ref_lib, tmp_path, name):
    """Assignment 9 at its default sampling, rays_per_pixel = 100 (10 x 10 lens grid, A09/code.js:232-235), on the reference's
    own scenes: launcher sequence and the one-launch frame against the reference kernels, bit for bit."""
    fx = G.load(name)
    P = fx["params"]
    cols, rows, rpp, n_slabs = 160, 120, 100, 5
    path = G.materialize_scene(fx["tree"], [], tmp_path)
    o_scene, p_scene = OH.loadScene(path, cols, rows, assignment=9), rt.loadScene(path, cols, rows, assignment=9)
    acu_o, pix_o, st = OR.a09_render(ref_lib, o_scene, cols, rows, rpp, n_slabs)
    acu, pix, matid, maxt = rt.assignments.a09_render(gpu_ctx, p_scene, cols, rows, rpp, n_slabs)
    want_id = st["pois"]["matId"].astype(np.int32)
    assert (matid == want_id).mean() >= 0.9999 and np.array_equal(matid, want_id), "hit ids"
    hit = want_id >= 0
    assert hit.any()
    assert np.array_equal(maxt[hit].view(np.uint32), st["rays"]["maxt"][hit].view(np.uint32))
    acu_o = np.ascontiguousarray(acu_o, dtype=np.float32)
    assert np.abs(acu - acu_o).max() <= 1e-4
    assert np.array_equal(acu.view(np.uint32), acu_o.view(np.uint32)), "expected bit-exact float image"
    assert np.array_equal(pix.reshape(-1, 4), np.asarray(pix_o).reshape(-1, 4))
    acu_f, pix_f, matid_f, _ = rt.assignments.a089_render_fused(gpu_ctx, p_scene, cols, rows, 9, rpp, n_slabs)
    assert np.array_equal(acu_f.view(np.uint32), acu.view(np.uint32)) and np.array_equal(matid_f, matid) and np.array_equal(pix_f, pix)
    assert P["assignment"] == 9


def _probe_rays(rng, bmin, bmax, n):
    """Random and adversarial rays around a set's box (o, d float32; not normalised on purpose for some)."""
    c, ext = (bmin + bmax) / 2, (bmax - bmin)
    o = np.empty((n, 3), np.float64)
    d = np.empty((n, 3), np.float64)
    k = n // 8
    # (a) from a shell around the box towards random points in / near it
    u = rng.normal(size=(3 * k, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    o[:3 * k] = c + u * ext.max() * rng.uniform(0.9, 3.0, size=(3 * k, 1))
    d[:3 * k] = (c + rng.uniform(-0.8, 0.8, size=(3 * k, 3)) * ext) - o[:3 * k]
    # (b) origins inside the box, random directions
    o[3 * k:5 * k] = bmin + rng.uniform(0, 1, size=(2 * k, 3)) * ext
    d[3 * k:5 * k] = rng.normal(size=(2 * k, 3))
    # (c) far origins (large entry parameters: the reference's t_next sequences drift)
    u = rng.normal(size=(k, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    o[5 * k:6 * k] = c + u * ext.max() * rng.choice([30.0, 1e3, 1e5], size=(k, 1))
    d[5 * k:6 * k] = (c + rng.uniform(-0.7, 0.7, size=(k, 3)) * ext) - o[5 * k:6 * k]
    # (d) axis-parallel and nearly axis-parallel directions (zero / tiny components)
    o[6 * k:7 * k] = c + rng.uniform(-1.5, 1.5, size=(k, 3)) * ext
    dd = rng.normal(size=(k, 3))
    dd[np.arange(k), rng.integers(0, 3, k)] *= rng.choice([0.0, 1e-30, 1e-12, 1e-6], size=k)
    d[6 * k:7 * k] = dd
    # (e) grazing: along the box faces and the surface-hugging corners
    face = rng.integers(0, 3, n - 7 * k)
    oo = bmin + rng.uniform(0, 1, size=(n - 7 * k, 3)) * ext
    oo[np.arange(len(face)), face] = np.where(rng.random(len(face)) < 0.5, bmin[face], bmax[face]) + rng.choice([0.0, 1e-6, -1e-6, 1e-3], size=len(face)) * ext[face]
    oo -= 2.0 * ext * rng.normal(size=(len(face), 1)) * 0  # stay on the face
    dg = rng.normal(size=(len(face), 3))
    dg[np.arange(len(face)), face] *= rng.choice([0.0, 1e-9, 1e-4, 0.05], size=len(face))
    o[7 * k:] = oo - dg * rng.uniform(0.0, 2.0, size=(len(face), 1))
    d[7 * k:] = dg
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    unit = rng.random(n) < 0.8
    d[unit] = d[unit] / np.where(nrm[unit] > 0, nrm[unit], 1.0)
    rays = np.zeros(n, OR.RAY)
    rays["o"][:, :3] = o.astype(np.float32)
    rays["d"][:, :3] = d.astype(np.float32)
    rays["mint"] = 0.0
    rays["maxt"] = np.inf
    return rays


@pytest.mark.parametrize("mesh_uv,nslabs", [((200, 100), 64), ((160, 80), 128), ((40, 20), 10)])
def test_skipped_walks_cross_empty_cells_only(rt, instr_lib, tmp_path, mesh_uv, nslabs):
    """The wavefront path does not queue a ray whose walk it can PROVE crosses empty cells only (walkProvablyEmpty, a march through
    a distance field of the coarse occupancy).  The proof must be conservative: every ray it flags must, in the reference's own
    kernel (instrumented build), visit cells but test NO reference -- on random rays, far origins (drifting t_next sequences),
    axis-parallel directions and rays grazing the box faces.  It must also be worth having: a fair share of the zero-test walks
    is flagged."""
    o_scene, p_scene = util.make_scene_pair(tmp_path, 64, 48, mesh_uv=mesh_uv, mesh_nslabs=nslabs)
    om = o_scene["meshes"][0]
    prep = OR.prepare_a10(o_scene, 1)
    mesh = prep["sets"][-1]
    bmin, bmax = np.array(om.bounds.min, np.float64), np.array(om.bounds.max, np.float64)
    n = 1 << 21
    rays = _probe_rays(np.random.Generator(np.random.PCG64(4242 + nslabs)), bmin, bmax, n)
    # the reference: meshTrace (closest hit) over the same rays with the counting hooks
    hit, cells, tests = np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    instr_lib.set_stats(hit, cells, tests)
    try:
        pois = np.zeros(n, OR.POI10)
        ro = rays.copy()
        instr_lib.a10_meshTrace(n, pois, ro, mesh["pos"], mesh["normal"], mesh["box"], mesh["matid"], mesh["aabb"], mesh["n"])
    finally:
        instr_lib.set_stats(None, None, None)
    r = rt.Renderer(p_scene, 64, 48, 4)
    r.preRender(OR.make_seeds(64 * 48 * 4, 1))
    try:
        ctx = r.ctx
        d_rays, d_flags = ctx.upload(rays), ctx.alloc(n)
        ctx.check(rt.lib.dll.rt_scene_probe_empty_walks(r.h_scene, len(prep["sets"]) - 1, d_rays, n, d_flags))
        flags = ctx.download(d_flags, np.uint8, n).astype(bool)
        ctx.free(d_rays)
        ctx.free(d_flags)
    finally:
        r.postRender()
    walked = cells > 0
    assert not (flags & ~walked).any(), "a ray that never enters the grid was flagged"
    bad = flags & (tests > 0)
    assert not bad.any(), "%d flagged rays test references in the reference kernel (first: %d, %d tests)" % (bad.sum(), np.flatnonzero(bad)[0], tests[bad][0])
    assert not (flags & (hit != NONE)).any()
    empty = walked & (tests == 0)
    share = flags.sum() / max(int(empty.sum()), 1)
    assert empty.sum() > 1000
    if nslabs >= 64:
        assert share > 0.3, "only %.1f %% of the zero-test walks are proven empty" % (100 * share)


def test_shadow_rays_skipped_against_walls_are_unblocked(rt, ref_lib, tmp_path):
    """The stage kernels do not test a shadow segment against a room's wall triangles when both its ends lie inside the room
    with rounding-proof margins (shadowClearsWalls).  Every segment the predicate flags must be left unblocked by the
    reference's own triangleShadowTrace -- on segments hugging the walls, ending at / beyond them, and nearly parallel to them."""
    o_scene, p_scene = util.make_scene_pair(tmp_path, 64, 48, mesh_uv=(24, 12), mesh_nslabs=8)
    prep = OR.prepare_a10(o_scene, 1)
    kinds = [q["kind"] for q in prep["sets"]]
    k = kinds.index("triangle")
    tri = prep["sets"][k]
    rng = np.random.Generator(np.random.PCG64(99))
    n = 1 << 21
    lo, hi = np.array([-0.98, -0.98, -0.98]), np.array([0.98, 0.98, 2.5])   # the synthetic room (tests/synth.py)
    o = lo + rng.uniform(0, 1, size=(n, 3)) * (hi - lo)
    # a third of the origins sit on a wall, pushed in by the renderer's 0.001 (or less, or outside)
    m = n // 3
    ax = rng.integers(0, 3, m)
    side = rng.random(m) < 0.5
    off = rng.choice([1e-3, 1e-3, 1e-4, 1e-5, 0.0, -1e-3], size=m)
    o[np.arange(m), ax] = np.where(side, lo[ax] + off, hi[ax] - off)
    q = lo + rng.uniform(-0.05, 1.05, size=(n, 3)) * (hi - lo)          # end points: inside, near the walls, some outside
    par = rng.random(n) < 0.2                                             # nearly parallel to a wall
    ax2 = rng.integers(0, 3, n)
    q[par, ax2[par]] = o[par, ax2[par]] + rng.choice([0.0, 1e-7, 1e-5, 1e-3], size=par.sum()) * rng.choice([-1, 1], size=par.sum())
    d = q - o
    L = np.linalg.norm(d, axis=1)
    keep = L > 1e-6
    o, d, L = o[keep], d[keep] / L[keep, None], L[keep]
    n = len(o)
    rays = np.zeros(n, OR.RAY)
    rays["o"][:, :3] = o.astype(np.float32)
    rays["d"][:, :3] = d.astype(np.float32)
    rays["d"][:, :3] /= np.linalg.norm(rays["d"][:, :3].astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    rays["mint"] = 0.0
    rays["maxt"] = L.astype(np.float32)
    after = rays.copy()
    ref_lib.a10_triangleShadowTrace(n, after, tri["pos"], tri["box"], tri["aabb"], tri["n"])
    blocked = after["mint"] == after["maxt"]
    r = rt.Renderer(p_scene, 64, 48, 4)
    r.preRender(OR.make_seeds(64 * 48 * 4, 1))
    try:
        ctx = r.ctx
        d_rays, d_flags = ctx.upload(rays), ctx.alloc(n)
        ctx.check(rt.lib.dll.rt_scene_probe_empty_walks(r.h_scene, k, d_rays, n, d_flags))
        flags = ctx.download(d_flags, np.uint8, n).astype(bool)
        ctx.free(d_rays)
        ctx.free(d_flags)
    finally:
        r.postRender()
    assert blocked.sum() > 1000 and (~blocked).sum() > 1000
    bad = flags & blocked
    assert not bad.any(), "%d flagged segments are blocked by a wall in the reference kernel (first: %d)" % (bad.sum(), np.flatnonzero(bad)[0])
    assert flags.mean() > 0.2, "only %.1f %% of the segments are skipped" % (100 * flags.mean())


@pytest.mark.parametrize("what", ["mesh", "mol"])
def test_a07_big_grids_through_the_queue_walker(rt, gpu_ctx, ref_lib, what):
    """Config 3 on big grids: molTrace / meshTrace of a library-built grid with >= 32 slabs and >= 64 K references run through the
    persistent pair-list walker (prepare -> k_walk_pairs with Assignment 7's exclusive triangle test -> finish) instead of one
    thread per ray.  Same hits, same hit distances, same debug colours as the reference kernels."""
    import synth
    cols, rows = 352, 264
    if what == "mesh":
        model = synth.synth_mesh(220, 110, seed=7)
        o_data, p_data, kw, n = OH.parseMeshJSON(model), rt.parseMeshJSON(model), "meshData", 64
    else:
        text = synth.synth_pdb(n_atoms=30000, gap_at=123)
        o_data, p_data, kw, n = OH.parsePDB(text), rt.parsePDB(text), "molData", 40
    g = (rt.splitMeshData if what == "mesh" else rt.splitMolData)(gpu_ctx, p_data, n)
    refs = int(g.n_refs)
    rt.lib.dll.rt_grid_release(gpu_ctx.h, C.byref(g))
    assert refs >= 65536, "the grid is too small to take the walker route (%d references)" % refs
    pix_o, rays_o, _ = OR.a07_render(ref_lib, cols, rows, n, **{kw: o_data})
    pix, maxt = rt.assignments.a07_compute(gpu_ctx, cols, rows, n, **{kw: p_data})
    assert np.array_equal(maxt.view(np.uint32), rays_o["maxt"].view(np.uint32)), "hit set / hit distances differ"
    assert np.array_equal(pix.reshape(-1, 4), np.asarray(pix_o).reshape(-1, 4)), "debug colours (shade x cell parity) differ"
    assert (pix.reshape(-1, 4)[:, :3].sum(axis=1) > 0).mean() > 0.05
