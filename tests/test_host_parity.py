"""CPU tests (no GPU): the product's host mirror of code.js (2015-raytracing_b200/host.py:
loaders, Bounds, Camera, Light packings, Mesh transforms) against the oracle's literal JS-order
restatement (oracle/host.py) on the golden demo inputs and on synthetic assets; and the
oracle's vectorised twins (used only for inputs too large for the literal loops) against the
literal versions."""
import numpy as np
import pytest

import golden_io as G
import synth
from oracle import host as OH


def _b(x):
    return (list(map(float, x.min)), list(map(float, x.max)))


@pytest.mark.parametrize("name", G.names("a10_"))
def test_loadScene_matches_oracle(rt, tmp_path, name):
    fx = G.load(name)
    P = fx["params"]
    path = G.materialize_scene(fx["tree"], G.meshes_of(fx), tmp_path)
    o = OH.loadScene(path, P["cols"], P["rows"], assignment=10)
    p = rt.loadScene(path, P["cols"], P["rows"])
    assert np.array_equal(p["camera"].toFloat32Array().view(np.uint32), o["camera"].toFloat32Array().view(np.uint32))
    assert np.array_equal(p["camera"].toFloat32Array().view(np.uint32), fx["cam16"].view(np.uint32))
    assert (p["focal_length"], p["lens_diameter"]) == (o["focal_length"], o["lens_diameter"])
    assert len(p["lights"]) == len(o["lights"])
    for a, b in zip(p["lights"], o["lights"]):
        for f in ("toShadowInfo", "toSceneRenderInfo", "toLightRenderInfo"):
            assert np.array_equal(getattr(a, f)().view(np.uint32), getattr(b, f)().view(np.uint32)), f
    assert np.array_equal(rt.splitMaterialData(p).view(np.uint32), OH.splitMaterialData(o).view(np.uint32))
    for k in ("bounds", "sphereBounds", "triangleBounds"):
        assert _b(p[k]) == _b(o[k]), k
        assert np.array_equal(rt.bounds2AABB(p[k]).view(np.uint32), OH.bounds2AABB(o[k]).view(np.uint32))
    assert len(p["spheres"]) == len(o["spheres"]) and len(p["triangles"]) == len(o["triangles"])
    for a, b in zip(p["spheres"], o["spheres"]):
        assert (a["c"].tolist(), a["r"], a["matId"]) == ([b["c"].x, b["c"].y, b["c"].z], b["r"], b["matId"])
    for a, b in zip(p["triangles"], o["triangles"]):
        for f in ("p0", "p1", "p2", "n0", "n1", "n2"):
            assert a[f].tolist() == [b[f].x, b[f].y, b[f].z]
        assert a["matId"] == b["matId"]
    assert len(p["meshes"]) == len(o["meshes"])
    for a, b in zip(p["meshes"], o["meshes"]):
        assert (a.ntriangles, int(a.nslabs), a.matId) == (b.ntriangles, int(b.nslabs), b.matId)
        assert _b(a.bounds) == _b(b.bounds)      # after normalize / scale / translate


@pytest.mark.parametrize("name", G.names("mol_"))
def test_parsePDB_matches_oracle(rt, name):
    fx = G.load(name)
    text = G.pdb_text(fx["serial"], fx["elem"], fx["xyz"])
    o, p = OH.parsePDB(text), rt.parsePDB(text)
    assert p["size"] == o["size"] == fx["params"]["size"]
    assert list(p["atomData"]) == list(o["atomData"])
    assert list(p["colorData"]) == list(o["colorData"]) and list(p["radiusData"]) == list(o["radiusData"])
    assert _b(p["bounds"]) == _b(o["bounds"])


def test_parsePDB_serial_gap_quirk(rt):
    """Q13: a TER record consumes a serial, so `size` (largest serial) exceeds the record count."""
    text = synth.synth_pdb(n_atoms=80, gap_at=33)
    o, p = OH.parsePDB(text), rt.parsePDB(text)
    assert o["size"] == p["size"] == 81 and len(o["atomData"]) == len(p["atomData"]) == 80 * 4


def _mesh_model(n_u, n_v):
    m = synth.synth_mesh(n_u, n_v, model_matrix=[0.5, 0.1, 0, 0, -0.1, 0.7, 0.2, 0, 0, 0.3, 1.1, 0, 0.25, -1.0, 3.0, 1])
    return m


def test_parseMeshJSON_matches_oracle(rt, tmp_path):
    """Node transform with gl-matrix's Float32Array rounding points (A10/lib/gl-matrix.js:79-80)."""
    model = _mesh_model(14, 9)
    p = tmp_path / "m.json"
    synth.mesh_to_json_file(model, str(p))
    o, q = OH.parseMeshJSON(str(p)), rt.parseMeshJSON(str(p))
    assert q["nTriangles"] == o["nTriangles"] == 2 * 14 * 9
    assert np.array_equal(np.asarray(q["positions"]).reshape(-1), np.asarray(o["positions"], dtype=np.float64))
    assert np.array_equal(np.asarray(q["normals"]).reshape(-1), np.asarray(o["normals"], dtype=np.float64))
    assert _b(q["bounds"]) == _b(o["bounds"])
    assert list(q["materialIndices"]) == list(o["materialIndices"]) and list(q["materials"]) == list(o["materials"])


@pytest.mark.parametrize("name", G.names("tri_"))
def test_parseMeshJSON_golden_meshes(rt, tmp_path, name):
    fx = G.load(name)
    m = G.meshes_of(fx)[0]
    p = tmp_path / "m.json"
    p.write_text(G.mesh_json_text(m["positions"], m["normals"], m["materialIndices"], m["materials"]))
    o, q = OH.parseMeshJSON(str(p)), rt.parseMeshJSON(str(p))
    assert np.array_equal(np.asarray(q["positions"]).reshape(-1), np.asarray(o["positions"], dtype=np.float64))
    assert np.array_equal(np.asarray(q["positions"]).reshape(-1), m["positions"])
    assert _b(q["bounds"]) == _b(o["bounds"])
    assert list(q["materialIndices"]) == list(o["materialIndices"])


def test_camera_set_and_rotate_match_oracle(rt):
    """Camera.set / rotate of the molecule and mesh demos (A07/code.js:45-124)."""
    b_o, b_p = OH.Bounds([-3.0, -1.5, 0.25], [4.0, 2.5, 9.0]), rt.Bounds([-3.0, -1.5, 0.25], [4.0, 2.5, 9.0])
    co, cp = OH.Camera(), rt.Camera()
    co.defaultInit(); cp.defaultInit()
    co.set(b_o, 320, 240); cp.set(b_p, 320, 240)
    assert np.array_equal(co.toFloat32Array().view(np.uint32), cp.toFloat32Array().view(np.uint32))
    for ang in (0, 10, 95.5, 270):
        co.rotate(b_o, ang); cp.rotate(b_p, ang)
        assert np.array_equal(co.toFloat32Array().view(np.uint32), cp.toFloat32Array().view(np.uint32))


# ------------------------------------------------------------------- oracle: vectorised twins
@pytest.mark.parametrize("n", [1, 2, 7, 16])
def test_oracle_fast_grid_build_equals_literal(monkeypatch, n):
    model = _mesh_model(20, 12)
    monkeypatch.setattr(OH, "FAST_MIN_PRIMS", 10 ** 9)
    lit_mesh = OH.parseMeshJSON(model)
    lit = OH.splitMeshData(lit_mesh, n)
    monkeypatch.setattr(OH, "FAST_MIN_PRIMS", 0)
    fast_mesh = OH.parseMeshJSON(model)
    fast = OH.splitMeshData(fast_mesh, n)
    assert np.array_equal(np.asarray(lit_mesh["positions"], dtype=np.float64), fast_mesh["positions"])
    assert np.array_equal(np.asarray(lit_mesh["normals"], dtype=np.float64), fast_mesh["normals"])
    assert _b(lit_mesh["bounds"]) == _b(fast_mesh["bounds"])
    for a, b in zip(lit, fast):
        assert np.array_equal(a, b)


def test_oracle_fast_grid_build_drops_and_clamps_like_the_loop(monkeypatch):
    """Primitives on the max face are dropped (lo == n is not clamped, the loop is empty), ones
    outside are clamped one-sidedly, NaN boxes land nowhere -- A10/code.js:940-965."""
    pos = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0],             # inside
                    [4, 4, 4, 4, 4, 4, 4, 4, 4],             # degenerate, on the max corner -> dropped
                    [-9, -9, -9, -8, -8, -8, -7, -7, -7],    # outside below -> hi < 0 ... loop from 0 to hi<0 empty
                    [3.5, 3.5, 3.5, 9, 9, 9, 3.6, 3.6, 3.6],  # sticks out above -> clamped
                    [np.nan] * 9], dtype=np.float64)
    md = {"positions": pos.reshape(-1), "normals": np.zeros(pos.size), "materialIndices": np.arange(len(pos)),
          "bounds": OH.Bounds([0.0, 0.0, 0.0], [4.0, 4.0, 4.0])}
    monkeypatch.setattr(OH, "FAST_MIN_PRIMS", 10 ** 9)
    lit = OH.splitMeshData(md, 4)
    monkeypatch.setattr(OH, "FAST_MIN_PRIMS", 0)
    fast = OH.splitMeshData(md, 4)
    for a, b in zip(lit, fast):
        assert np.array_equal(a, b, equal_nan=True)
    assert sorted(set(lit[3].tolist())) == [0, 3]


# ------------------------------------------------------------------- native loaders (rt_parse_mesh_json / rt_parse_pdb)
def _same_mesh(n, o):
    assert n["nTriangles"] == o["nTriangles"] and n["nMaterials"] == o["nMaterials"]
    assert np.array_equal(np.asarray(n["positions"]).reshape(-1), np.asarray(o["positions"], dtype=np.float64))
    assert np.array_equal(np.asarray(n["normals"]).reshape(-1), np.asarray(o["normals"], dtype=np.float64))
    assert list(n["materialIndices"]) == list(o["materialIndices"]) and list(n["materials"]) == list(o["materials"])
    assert _b(n["bounds"]) == _b(o["bounds"])


def test_native_mesh_loader_matches_oracle(rt, tmp_path):
    """Node transform + fp32 rounding points, indexed mesh, unknown keys, nested values, BOM."""
    model = _mesh_model(14, 9)
    p = tmp_path / "m.json"
    synth.mesh_to_json_file(model, str(p))
    _same_mesh(rt.parseMeshJSON_native(str(p)), OH.parseMeshJSON(str(p)))
    text = '\ufeff{"name": "x", "extra": {"a": [1, {"b": null}, "s\\"q"], "t": true}, ' + p.read_text()[1:]
    q = tmp_path / "m2.json"
    q.write_text(text, encoding="utf-8")
    _same_mesh(rt.parseMeshJSON_native(str(q)), OH.parseMeshJSON(str(p)))


@pytest.mark.parametrize("name", G.names("tri_"))
def test_native_mesh_loader_golden_meshes(rt, tmp_path, name):
    fx = G.load(name)
    m = G.meshes_of(fx)[0]
    p = tmp_path / "m.json"
    p.write_text(G.mesh_json_text(m["positions"], m["normals"], m["materialIndices"], m["materials"]))
    _same_mesh(rt.parseMeshJSON_native(str(p)), OH.parseMeshJSON(str(p)))


def test_native_mesh_loader_rejects_malformed(rt):
    for bad in (b"", b"{", b'{"meshes": [{"vertexPositions": [1, 2,', b'{"materials": [], "meshes": [], "nodes": [{"modelMatrix": [1, 2], "meshIndices": []}]}',
                b'{"materials": [], "meshes": [], "nodes": [{"modelMatrix": [0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0], "meshIndices": []}]}'):
        with pytest.raises(ValueError):
            rt.parseMeshJSON_native(bad)


@pytest.mark.parametrize("name", G.names("mol_"))
def test_native_pdb_loader_matches_oracle(rt, name):
    fx = G.load(name)
    text = G.pdb_text(fx["serial"], fx["elem"], fx["xyz"])
    o, n = OH.parsePDB(text), rt.parsePDB_native(text)
    assert n["size"] == o["size"] and list(n["atomData"]) == list(o["atomData"])
    assert list(n["colorData"]) == list(o["colorData"]) and list(n["radiusData"]) == list(o["radiusData"])
    assert _b(n["bounds"]) == _b(o["bounds"])


def test_native_pdb_loader_serial_gap_and_altloc(rt):
    text = synth.synth_pdb(n_atoms=80, gap_at=33)
    lines = text.split("\n")
    lines.insert(5, lines[4][:16] + "B" + lines[4][17:])     # an altLoc 'B' duplicate: skipped by both
    text = "\n".join(lines)
    o, n = OH.parsePDB(text), rt.parsePDB_native(text)
    assert n["size"] == o["size"] == 81 and list(n["atomData"]) == list(o["atomData"])


def test_native_json_numbers_are_correctly_rounded(rt):
    """The parser's exact fast path (significand < 2^53, |exp10| <= 22) and its strtod fallback against Python's
    float() (correctly rounded, like JavaScript's number parsing), compared as raw doubles."""
    import ctypes as C
    import random
    random.seed(3)
    toks = []
    for _ in range(60000):
        k = random.random()
        if k < 0.3:
            toks.append(repr(random.uniform(-1e3, 1e3)))
        elif k < 0.45:
            toks.append("%.6f" % random.uniform(-10, 10))
        elif k < 0.6:
            toks.append("%de%d" % (random.randint(-10 ** 15, 10 ** 15), random.randint(-40, 40)))
        elif k < 0.7:
            toks.append("%.17g" % random.uniform(-1, 1))
        elif k < 0.8:
            toks.append(str(random.randint(-2 ** 62, 2 ** 62)))
        elif k < 0.9:
            toks.append("%.3e" % random.uniform(-1e30, 1e30))
        else:
            toks.append(random.choice(["0", "-0", "-0.0", "1e22", "1e23", "9007199254740993", "9007199254740992", "0.1", "1E5",
                                       "12345678901234567890123", "-0.000001", "5e-324", "1.7976931348623157e308", "2.2250738585072014e-308",
                                       "1e-22", "1e-23", "123456789012345678e-25"]))
    n = len(toks) - len(toks) % 4
    toks = toks[:n]
    mats = ",".join('{"diffuseReflectance":[%s]}' % ",".join(toks[i:i + 4]) for i in range(0, n, 4))
    raw = ('{"materials":[%s],"meshes":[]}' % mats).encode()
    md = rt.lib.MeshData()
    assert rt.lib.dll.rt_parse_mesh_json(raw, len(raw), C.byref(md)) == 0
    try:
        got = np.ctypeslib.as_array(md.materials, shape=(n,)).copy()
    finally:
        rt.lib.dll.rt_mesh_data_free(C.byref(md))
    want = np.array([float(t) for t in toks])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
