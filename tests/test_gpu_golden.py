"""GPU parity tests against the committed golden fixtures (tests/golden/, produced by the oracle
from the reference's own demo inputs -- see tests/golden/make_golden.py): every BASELINE.json
config family through the C ABI, no oracle and no /root/reference needed at run time.

Gates: grid cell lists bit-exact; hit ids >= 99.99 % equal; deterministic
float images within 1e-4 per channel; path-traced images within 1e-3 per channel at equal spp with
the seed buffer equal as integers.  In practice everything is bit-exact and asserted as such
where the arithmetic contract makes that the expected outcome."""
import ctypes as C

import numpy as np
import pytest

import golden_io as G

pytestmark = pytest.mark.gpu


def _grid_digests(rt, ctx, g, cells=None):
    d = rt.host.DeviceGrid(ctx, g, np.zeros(8, np.float32), cells=cells)
    out = {"refs": d.n_refs, "box": G.digest(d.box_size()), "prim": G.digest(d.prim())}
    if g.kind == 1 and g.normal:
        out["normal"] = G.digest(d.normal())
    if g.matid:
        out["matid"] = G.digest(d.matid())
    return out


def _same_pixels(got, want, what, budget=1e-4):
    got, want = np.asarray(got).reshape(-1, 4), np.asarray(want).reshape(-1, 4)
    bad = np.any(got != want, axis=1).mean()
    assert bad <= budget, "%s: %.4f%% of the pixels differ" % (what, 100 * bad)


def test_a01(rt, gpu_ctx):
    fx = G.load("a01")
    for cols, rows in fx["params"]["sizes"]:
        pix = rt.assignments.a01_compute(gpu_ctx, cols, rows)
        assert np.array_equal(pix, fx["pixels_%dx%d" % (cols, rows)])


def _same_bits(got, want, what):
    assert np.array_equal(np.asarray(got).view(np.uint32), np.asarray(want).view(np.uint32)), what


def _check_a456(rt, ctx, fx, cols, rows, slabs, mol=None, mesh=None, prefix="", slab_grids=None):
    """A04 / A05 (brute force, + boxes) and A06 (x slabs built on the GPU) frames, SURVEY.md 8f rank 4."""
    A = rt.assignments
    p, maxt = A.a04_compute(ctx, cols, rows, molData=mol, meshData=mesh)
    _same_bits(maxt, fx[prefix + "a04_maxt"], "A04 hit distance / hit set")
    assert np.array_equal(p, fx[prefix + "a04_pixels"]), "A04 pixels"
    p, maxt = A.a05_compute(ctx, cols, rows, molData=mol, meshData=mesh)
    _same_bits(maxt, fx[prefix + "a05_maxt"], "A05 hit distance / hit set")
    assert np.array_equal(p, fx[prefix + "a05_pixels"]), "A05 pixels"
    if mol is not None and mesh is None:
        assert np.array_equal(A.a04_raytrace(ctx, mol, cols, rows), fx[prefix + "a04_raytrace"]), "A04 raytrace"
    for k, n in enumerate(slabs):
        if slab_grids is not None:
            want = slab_grids[k]
            if "mol" in want:
                g = rt.slabSplitMolData(ctx, mol, n)
                got = _grid_digests(rt, ctx, g, cells=n)
                rt.lib.dll.rt_grid_release(ctx.h, C.byref(g))
                assert (got["refs"], got["box"], got["prim"], got["matid"]) == (
                    want["mol"]["refs"], want["mol"]["box"], want["mol"]["prim"], want["mol"]["index"]), "A06 atom slabs differ at n=%d" % n
            if "mesh" in want:
                g = rt.slabSplitMeshData(ctx, mesh, n)
                got = _grid_digests(rt, ctx, g, cells=n)
                rt.lib.dll.rt_grid_release(ctx.h, C.byref(g))
                assert (got["refs"], got["box"], got["prim"], got["normal"], got["matid"]) == (
                    want["mesh"]["refs"], want["mesh"]["box"], want["mesh"]["prim"], want["mesh"]["normal"], want["mesh"]["index"]), \
                    "A06 triangle slabs differ at n=%d" % n
        p, maxt = A.a06_compute(ctx, cols, rows, n, molData=mol, meshData=mesh)
        _same_bits(maxt, fx[prefix + "a06_maxt_n%d" % n], "A06 hit distance / hit set n=%d" % n)
        assert np.array_equal(p, fx[prefix + "a06_pixels_n%d" % n]), "A06 pixels n=%d" % n


@pytest.mark.parametrize("name", G.names("mol_"))
def test_molecule(rt, gpu_ctx, name):
    """A02 fused kernel, A03 two kernels (ray mint = hit distance), A07 grid + molTrace."""
    fx = G.load(name)
    P = fx["params"]
    mol = rt.parsePDB(G.pdb_text(fx["serial"], fx["elem"], fx["xyz"]))
    assert mol["size"] == P["size"]
    cols, rows = P["cols"], P["rows"]
    assert np.array_equal(rt.assignments.a02_compute(gpu_ctx, mol, cols, rows), fx["a02_pixels"])
    p3, mint = rt.assignments.a03_compute(gpu_ctx, mol, cols, rows)
    assert np.array_equal(p3, fx["a03_pixels"])
    assert np.array_equal(mint.view(np.uint32), fx["a03_mint"].view(np.uint32))
    for n, want in zip(P["slabs"], fx["grids"]):
        g = rt.splitMolData(gpu_ctx, mol, n)
        got = _grid_digests(rt, gpu_ctx, g)
        rt.lib.dll.rt_grid_release(gpu_ctx.h, C.byref(g))
        assert (got["refs"], got["box"], got["prim"], got["matid"]) == (want["refs"], want["box"], want["prim"], want["index"])
        p7, maxt = rt.assignments.a07_compute(gpu_ctx, cols, rows, n, molData=mol)
        assert np.array_equal(maxt.view(np.uint32), fx["a07_maxt_n%d" % n].view(np.uint32)), "hit distance / hit set differs"
        assert np.array_equal(p7, fx["a07_pixels_n%d" % n])
    _check_a456(rt, gpu_ctx, fx, cols, rows, P["slabs"], mol=mol, slab_grids=fx["slab_grids"])


@pytest.mark.parametrize("name", G.names("tri_"))
def test_mesh(rt, gpu_ctx, tmp_path, name):
    """Grid build at several resolutions (bit-exact cell lists) and A07 meshTrace (exclusive
    triangle range test, quirk Q9); teapot also runs computeBoth with a molecule."""
    fx = G.load(name)
    P = fx["params"]
    m = G.meshes_of(fx)[0]
    p = tmp_path / "m.json"
    p.write_text(G.mesh_json_text(m["positions"], m["normals"], m["materialIndices"], m["materials"]))
    md = rt.parseMeshJSON(str(p))
    for want in fx["grids"]:
        g = rt.splitMeshData(gpu_ctx, md, want["n"])
        got = _grid_digests(rt, gpu_ctx, g)
        rt.lib.dll.rt_grid_release(gpu_ctx.h, C.byref(g))
        assert (got["refs"], got["box"], got["prim"], got["normal"], got["matid"]) == (
            want["refs"], want["box"], want["prim"], want["normal"], want["index"]), "cell lists differ at n=%d" % want["n"]
    for n in P["slabs"]:
        p7, maxt = rt.assignments.a07_compute(gpu_ctx, P["cols"], P["rows"], n, meshData=md)
        assert np.array_equal(maxt.view(np.uint32), fx["a07_maxt_n%d" % n].view(np.uint32))
        assert np.array_equal(p7, fx["a07_pixels_n%d" % n])
    _check_a456(rt, gpu_ctx, fx, P["cols"], P["rows"], P["slabs"], mesh=md, slab_grids=fx["slab_grids"])
    if P.get("with_mol"):
        mol = rt.parsePDB(G.pdb_text(fx["both_serial"], fx["both_elem"], fx["both_xyz"]))
        pb, maxt = rt.assignments.a07_compute(gpu_ctx, P["cols"], P["rows"], 5, molData=mol, meshData=md)
        assert np.array_equal(maxt.view(np.uint32), fx["both_maxt"].view(np.uint32))
        assert np.array_equal(pb, fx["both_pixels"])
        _check_a456(rt, gpu_ctx, fx, P["cols"], P["rows"], (5,), mol=mol, mesh=md, prefix="both_")


@pytest.mark.parametrize("name", G.names("a08_") + G.names("a09_"))
def test_a08_a09_scene(rt, gpu_ctx, tmp_path, name):
    fx = G.load(name)
    P = fx["params"]
    scene = rt.loadScene(G.materialize_scene(fx["tree"], [], tmp_path), P["cols"], P["rows"], assignment=P["assignment"])
    assert np.array_equal(scene["camera"].toFloat32Array().view(np.uint32), fx["cam16"].view(np.uint32))
    if P["assignment"] == 8:
        acu, pix, matid, maxt = rt.assignments.a08_render(gpu_ctx, scene, P["cols"], P["rows"], P["n_slabs"])
    else:
        acu, pix, matid, maxt = rt.assignments.a09_render(gpu_ctx, scene, P["cols"], P["rows"], P["rpp"], P["n_slabs"])
    assert (matid == fx["matid"]).mean() >= 0.9999, "hit ids"
    assert np.abs(acu - fx["acu"]).max() <= 1e-4, "deterministic float image"
    assert np.array_equal(matid, fx["matid"]) and np.array_equal(acu.view(np.uint32), fx["acu"].view(np.uint32)), "expected bit-exact"
    hit = fx["matid"] >= 0
    assert np.array_equal(maxt[hit].view(np.uint32), fx["maxt"][hit].view(np.uint32))
    _same_pixels(pix, fx["pixels"], "uchar image")
    # the whole frame in one launch (rt_a089_render_frame) leaves the same accumulators, hit records and pixels
    acu_f, pix_f, matid_f, maxt_f = rt.assignments.a089_render_fused(gpu_ctx, scene, P["cols"], P["rows"], P["assignment"], P["rpp"], P["n_slabs"])
    assert np.array_equal(acu_f.view(np.uint32), acu.view(np.uint32)) and np.array_equal(matid_f, matid) and np.array_equal(pix_f, pix)
    assert np.array_equal(maxt_f[hit].view(np.uint32), maxt[hit].view(np.uint32))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", G.names("a10_"))
def test_a10_scene(rt, tmp_path, name, mode):
    """The reference's own A10 scenes (open scenes exercise quirk Q2, cornell_teapot3 the
    un-normalised light normal Q4 and two <mesh> grids): two progressive passes."""
    fx = G.load(name)
    P = fx["params"]
    scene = rt.loadScene(G.materialize_scene(fx["tree"], G.meshes_of(fx), tmp_path), P["cols"], P["rows"])
    total = P["cols"] * P["rows"] * P["rpp"]
    seeds = np.random.Generator(np.random.PCG64(P["seed"])).integers(1, 2 ** 31, size=total, dtype=np.int64).astype(np.int32)
    r = rt.Renderer(scene, P["cols"], P["rows"], P["rpp"], mode=mode)
    r.preRender(seeds)
    try:
        grids = list(r._grids) + [m.grid for m in scene["meshes"]]
        for g, want in zip(grids, fx["grids"]):
            got = _grid_digests(rt, r.ctx, g)
            assert (got["refs"], got["box"], got["prim"]) == (want["refs"], want["box"], want["prim"]), "cell lists of a %s set" % want["kind"]
        closest = anyh = 0
        for p in range(P["passes"]):
            pix = r.executeRender()
            acc = r.accum()
            assert np.abs(acc[:, :3] - fx["accum"][p][:, :3]).max() / (P["rpp"] * (p + 1)) <= 1e-3
            assert np.array_equal(r.seeds(), fx["seeds_after"][p]), "RNG streams (draw count/order) differ"
            assert np.array_equal(acc.view(np.uint32), fx["accum"][p].view(np.uint32)), "expected bit-exact accumulation"
            assert np.array_equal(pix, fx["pixels"][p])
            s = r.stats()
            closest += s["closest_rays"]
            anyh += s["any_rays"]
            assert [closest, anyh] == list(fx["counts"][p]), "valid-ray counts (the Mrays/s numerator)"
    finally:
        r.postRender()
