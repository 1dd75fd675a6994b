"""CPU tests (no GPU): the oracle -- the reference's kernels compiled from where they lie
(oracle/_ref) or the plain-C restatement, driven by the JS-order host restatement -- must
reproduce every golden fixture under tests/golden/ bit for bit.  The fixtures were produced by
tests/golden/make_golden.py from the reference's own demo inputs; this pins the oracle build
(compiler, flags, OpenMP schedule independence) on whatever machine runs the parity tests."""
import numpy as np
import pytest

import golden_io as G
from oracle import host as OH
from oracle import refcl as OR


def _eq(a, b, what):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, what
    assert a.tobytes() == b.tobytes(), what + ": differs from the golden fixture"


@pytest.mark.parametrize("name", G.names("a10_"))
def test_a10_scene(oracle_lib, tmp_path, name):
    fx = G.load(name)
    P = fx["params"]
    path = G.materialize_scene(fx["tree"], G.meshes_of(fx), tmp_path)
    scene = OH.loadScene(path, P["cols"], P["rows"], assignment=10)
    _eq(scene["camera"].toFloat32Array(), fx["cam16"], "camera float16")
    prep = OR.prepare_a10(scene, 1)
    for s, g in zip(prep["sets"], fx["grids"]):
        assert int(s["box"][-1]) == g["refs"] and G.digest(s["box"].astype(np.uint32)) == g["box"]
        assert G.digest(s["data"] if s["kind"] == "sphere" else s["pos"]) == g["prim"]
    total = P["cols"] * P["rows"] * P["rpp"]
    st = OR.A10State(total, OR.make_seeds(total, P["seed"]))
    oracle_lib.a10_initAcu(st.acu, total)
    for p in range(P["passes"]):
        pix = OR.a10_execute_render(oracle_lib, st, prep, fx["cam16"], P["cols"], P["rows"], P["rpp"], scene["focal_length"],
                                    scene["lens_diameter"])
        acc = np.zeros((P["cols"] * P["rows"], 4), np.float32)
        for k in range(P["rpp"]):
            acc += st.acu.reshape(-1, P["rpp"], 4)[:, k]
        _eq(acc, fx["accum"][p], "accumulation image, pass %d" % p)
        _eq(st.seeds, fx["seeds_after"][p], "seed buffer after pass %d" % p)
        _eq(pix, fx["pixels"][p], "pixels, pass %d" % p)
        assert [st.n_closest, st.n_any] == list(fx["counts"][p])


@pytest.mark.parametrize("name", G.names("a08_") + G.names("a09_"))
def test_a08_a09_scene(ref_lib, tmp_path, name):
    fx = G.load(name)
    P = fx["params"]
    path = G.materialize_scene(fx["tree"], [], tmp_path)
    scene = OH.loadScene(path, P["cols"], P["rows"], assignment=P["assignment"])
    if P["assignment"] == 8:
        acu, pix, st = OR.a08_render(ref_lib, scene, P["cols"], P["rows"], P["n_slabs"])
    else:
        acu, pix, st = OR.a09_render(ref_lib, scene, P["cols"], P["rows"], P["rpp"], P["n_slabs"])
    _eq(acu, fx["acu"], "acu")
    _eq(pix, fx["pixels"], "pixels")
    _eq(st["pois"]["matId"].astype(np.int32), fx["matid"], "hit material ids")
    _eq(st["rays"]["maxt"], fx["maxt"], "ray maxt")


@pytest.mark.parametrize("name", G.names("mol_"))
def test_molecule(ref_lib, name):
    fx = G.load(name)
    P = fx["params"]
    mol = OH.parsePDB(G.pdb_text(fx["serial"], fx["elem"], fx["xyz"]))
    assert mol["size"] == P["size"]
    _eq(OR.a02_render(ref_lib, mol, P["cols"], P["rows"]), fx["a02_pixels"], "A02 pixels")
    p3, r3 = OR.a03_render(ref_lib, mol, P["cols"], P["rows"])
    _eq(p3, fx["a03_pixels"], "A03 pixels")
    _eq(r3["mint"], fx["a03_mint"], "A03 ray mint")
    for n, g in zip(P["slabs"], fx["grids"]):
        p7, r7, prep = OR.a07_render(ref_lib, P["cols"], P["rows"], n, molData=mol)
        m = prep["mol"]
        assert (int(m["box"][-1]), G.digest(m["box"]), G.digest(m["atoms"]), G.digest(m["index"])) == (g["refs"], g["box"], g["prim"], g["index"])
        _eq(p7, fx["a07_pixels_n%d" % n], "A07 molTrace pixels n=%d" % n)
        _eq(r7["maxt"], fx["a07_maxt_n%d" % n], "A07 molTrace maxt n=%d" % n)
    _check_a456(ref_lib, fx, P["cols"], P["rows"], P["slabs"], mol=mol, slab_grids=fx["slab_grids"])


def _check_a456(lib, fx, cols, rows, slabs, mol=None, mesh=None, prefix="", slab_grids=None):
    """A04 / A05 / A06 frames against the fixture entries make_golden.a456_outputs wrote."""
    p, r = OR.a04_render(lib, cols, rows, molData=mol, meshData=mesh)
    _eq(p, fx[prefix + "a04_pixels"], "A04 pixels")
    _eq(r["maxt"], fx[prefix + "a04_maxt"], "A04 ray maxt")
    p, r = OR.a05_render(lib, cols, rows, molData=mol, meshData=mesh)
    _eq(p, fx[prefix + "a05_pixels"], "A05 pixels")
    _eq(r["maxt"], fx[prefix + "a05_maxt"], "A05 ray maxt")
    if mol is not None and mesh is None:
        _eq(OR.a04_raytrace(lib, mol, cols, rows), fx[prefix + "a04_raytrace"], "A04 raytrace")
    for k, n in enumerate(slabs):
        p, r, prep = OR.a06_render(lib, cols, rows, n, molData=mol, meshData=mesh)
        _eq(p, fx[prefix + "a06_pixels_n%d" % n], "A06 pixels n=%d" % n)
        _eq(r["maxt"], fx[prefix + "a06_maxt_n%d" % n], "A06 ray maxt n=%d" % n)
        if slab_grids is not None:
            g = slab_grids[k]
            if "mol" in g:
                m = prep["mol"]
                assert (int(m["box"][-1]), G.digest(m["box"]), G.digest(m["atoms"]), G.digest(m["colors"]), G.digest(m["index"])) == (
                    g["mol"]["refs"], g["mol"]["box"], g["mol"]["prim"], g["mol"]["colors"], g["mol"]["index"])
            if "mesh" in g:
                t = prep["mesh"]
                assert (int(t["box"][-1]), G.digest(t["box"]), G.digest(t["pos"]), G.digest(t["normal"]), G.digest(t["index"])) == (
                    g["mesh"]["refs"], g["mesh"]["box"], g["mesh"]["prim"], g["mesh"]["normal"], g["mesh"]["index"])


@pytest.mark.parametrize("name", G.names("tri_"))
def test_mesh(ref_lib, tmp_path, name):
    fx = G.load(name)
    P = fx["params"]
    m = G.meshes_of(fx)[0]
    p = tmp_path / "m.json"
    p.write_text(G.mesh_json_text(m["positions"], m["normals"], m["materialIndices"], m["materials"]))
    md = OH.parseMeshJSON(str(p))
    assert np.array_equal(np.asarray(md["positions"]), m["positions"])
    for g in fx["grids"]:
        pos, nor, box, idx = OH.splitMeshData(md, g["n"])
        assert (int(box[-1]), G.digest(box.astype(np.uint32)), G.digest(OH.to_f32(pos)), G.digest(OH.to_f32(nor))) == (
            g["refs"], g["box"], g["prim"], g["normal"])
    for n in P["slabs"]:
        p7, r7, _ = OR.a07_render(ref_lib, P["cols"], P["rows"], n, meshData=md)
        _eq(p7, fx["a07_pixels_n%d" % n], "A07 meshTrace pixels n=%d" % n)
        _eq(r7["maxt"], fx["a07_maxt_n%d" % n], "A07 meshTrace maxt n=%d" % n)
    _check_a456(ref_lib, fx, P["cols"], P["rows"], P["slabs"], mesh=md, slab_grids=fx["slab_grids"])
    if P.get("with_mol"):
        mol = OH.parsePDB(G.pdb_text(fx["both_serial"], fx["both_elem"], fx["both_xyz"]))
        pb, rb, _ = OR.a07_render(ref_lib, P["cols"], P["rows"], 5, molData=mol, meshData=md)
        _eq(pb, fx["both_pixels"], "A07 computeBoth pixels")
        _eq(rb["maxt"], fx["both_maxt"], "A07 computeBoth maxt")
        _check_a456(ref_lib, fx, P["cols"], P["rows"], (5,), mol=mol, mesh=md, prefix="both_")


def test_a01(ref_lib):
    fx = G.load("a01")
    for cols, rows in fx["params"]["sizes"]:
        _eq(OR.a01_render(ref_lib, cols, rows), fx["pixels_%dx%d" % (cols, rows)], "A01 %dx%d" % (cols, rows))


def test_struct_size_probes(ref_lib):
    """sizeofRay / sizeofPoi (A10/code.cl:440-446): the layouts every buffer is sized by."""
    assert ref_lib.a10_sizeofRay() == 48 and ref_lib.a10_sizeofPoi() == 64
    assert ref_lib.a08_sizeofRay() == 48 and ref_lib.a08_sizeofPoi() == 48
    assert ref_lib.a09_sizeofRay() == 48 and ref_lib.a09_sizeofPoi() == 48
    assert ref_lib.a03_sizeofRay() == 48 and ref_lib.a07_sizeofRay() == 48
    assert ref_lib.a04_sizeofRay() == 48 and ref_lib.a05_sizeofRay() == 48 and ref_lib.a06_sizeofRay() == 48


def test_instrumented_build_is_arithmetic_neutral(tmp_path):
    """The counting hooks of libref_instr.so sit at non-arithmetic places: same buffers as libref.so."""
    if not OR.have_reference():
        pytest.skip("oracle/_ref not built")
    fx = G.load("a10_cornell_teapot3")
    P = fx["params"]
    scene = OH.loadScene(G.materialize_scene(fx["tree"], G.meshes_of(fx), tmp_path), P["cols"], P["rows"])
    res = []
    for instr in (False, True):
        lib = OR.load_reference(instrumented=instr)
        assert lib.instrumented == instr
        st, pix, _ = OR.a10_render(lib, scene, P["cols"], P["rows"], P["rpp"], passes=1, seed=P["seed"])
        res.append((st.acu.copy(), st.seeds.copy(), pix.copy()))
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_instrumented_build_counts(tmp_path):
    """The hooks of libref_instr.so really fire (round 1 shipped an "instrumented" library that had been compiled from the
    un-hooked text and counted nothing): per pixel of an A07 meshTrace the winner champ_i is recorded exactly where the ray's
    maxt was lowered, every started walk visits at least one cell, tests only happen in visited cells, and the un-instrumented
    build leaves the sinks untouched."""
    if not OR.have_reference():
        pytest.skip("oracle/_ref not built")
    fx = G.load("tri_teapot")
    P = fx["params"]
    m = G.meshes_of(fx)[0]
    path = tmp_path / "m.json"
    path.write_text(G.mesh_json_text(m["positions"], m["normals"], m["materialIndices"], m["materials"]))
    md = OH.parseMeshJSON(str(path))
    cols, rows, n = P["cols"], P["rows"], 10
    out = {}
    for instr in (True, False):
        lib = OR.load_reference(instrumented=instr)
        hit = np.full(cols * rows, 7, np.uint32)
        cells, tests = np.full(cols * rows, 7, np.uint64), np.full(cols * rows, 7, np.uint64)
        lib.set_stats(hit, cells, tests)
        try:
            _, rays, prep = OR.a07_render(lib, cols, rows, n, meshData=md)
        finally:
            lib.set_stats(None, None, None)
        out[instr] = (hit, cells, tests, rays["maxt"].copy())
    hit, cells, tests, maxt = out[True]
    refs = int(prep["mesh"]["box"][-1])
    got = hit != 0xFFFFFFFF
    assert got.sum() > 100 and hit[got].max() < refs
    # a hit lowers maxt below the exit parameter initTrace stored; recompute initTrace to compare
    lib = OR.load_reference()
    pix0, rays0 = np.zeros((rows * cols, 4), np.uint8), np.zeros(rows * cols, dtype=OR.RAY)
    lib.a07_initTrace(pix0, prep["cam"], rays0, prep["aabb"], cols, rows)
    assert np.array_equal(got, maxt != rays0["maxt"])
    alive = rays0["mint"] != rays0["maxt"]
    assert np.array_equal(cells > 0, alive) or (cells[alive] > 0).all() and not cells[~alive].any()
    assert (tests[cells == 0] == 0).all() and tests.sum() > cells.sum() / 50 and tests[got].min() >= 1
    assert np.array_equal(out[False][3].view(np.uint32), maxt.view(np.uint32))
    assert (out[False][1] == 0).all() and (out[False][2] == 0).all() and (out[False][0] == 0xFFFFFFFF).all()


@pytest.mark.parametrize("name", G.names("a10_"))
def test_port_equals_reference_kernels(tmp_path, name):
    """The plain-C restatement of the Assignment-10 kernels (oracle/rt_oracle.c, kind "port") against the fixtures
    the reference's own kernel text produced: same accumulation image, seed buffer, pixels and ray counts, bit for bit."""
    port = OR.load_port()
    assert port.kind == "port"
    fx = G.load(name)
    P = fx["params"]
    scene = OH.loadScene(G.materialize_scene(fx["tree"], G.meshes_of(fx), tmp_path), P["cols"], P["rows"], assignment=10)
    prep = OR.prepare_a10(scene, 1)
    total = P["cols"] * P["rows"] * P["rpp"]
    st = OR.A10State(total, OR.make_seeds(total, P["seed"]))
    port.a10_initAcu(st.acu, total)
    for p in range(P["passes"]):
        pix = OR.a10_execute_render(port, st, prep, fx["cam16"], P["cols"], P["rows"], P["rpp"], scene["focal_length"], scene["lens_diameter"])
        acc = np.zeros((P["cols"] * P["rows"], 4), np.float32)
        for k in range(P["rpp"]):
            acc += st.acu.reshape(-1, P["rpp"], 4)[:, k]
        _eq(acc, fx["accum"][p], "accumulation image, pass %d" % p)
        _eq(st.seeds, fx["seeds_after"][p], "seed buffer after pass %d" % p)
        _eq(pix, fx["pixels"][p], "pixels, pass %d" % p)
        assert [st.n_closest, st.n_any] == list(fx["counts"][p])


def test_port_rpp1_and_row_tiles_equal_reference(tmp_path):
    """rays_per_pixel == 1 (serial seeds[col] order, Q7) and the row-tile entry of the port against the reference build."""
    if not OR.have_reference():
        pytest.skip("oracle/_ref not built")
    ref, port = OR.load_reference(), OR.load_port()
    fx = G.load("a10_cornell_teapot3")
    P = fx["params"]
    scene = OH.loadScene(G.materialize_scene(fx["tree"], G.meshes_of(fx), tmp_path), P["cols"], P["rows"], assignment=10)
    prep = OR.prepare_a10(scene, 1)
    cols, rows = P["cols"], P["rows"]
    out = []
    for lib in (ref, port):
        st = OR.A10State(cols * rows, OR.make_seeds(cols * rows, 3))
        lib.a10_initAcu(st.acu, st.total)
        pix = OR.a10_execute_render(lib, st, prep, fx["cam16"], cols, rows, 1, scene["focal_length"], scene["lens_diameter"], serial_init=True)
        st2 = OR.A10State(cols * 5 * 4, OR.make_seeds(cols * 5 * 4, 4))
        lib.a10_initAcu(st2.acu, st2.total)
        OR.a10_execute_render(lib, st2, prep, fx["cam16"], cols, rows, 4, scene["focal_length"], scene["lens_diameter"], row0=7, nrows=5)
        out.append((st.acu.copy(), st.seeds.copy(), pix.copy(), st2.acu.copy(), st2.seeds.copy()))
    for a, b in zip(*out):
        assert a.tobytes() == b.tobytes()
