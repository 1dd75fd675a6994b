#!/usr/bin/env python3
"""Generates tests/golden/*.npz -- run in the build container, where /root/reference exists and
oracle/_ref (the reference's own code.cl compiled by g++) has been built:

    python tests/golden/make_golden.py

For each reference demo input it (1) reduces the input to neutral numbers (golden_io.describe_*),
(2) re-emits it into a temp directory in the reference's on-disk formats, (3) runs the ORACLE on
the re-emitted files -- the reference's kernels (oracle/_ref/libref.so) driven by the JS-order
host restatement (oracle/host.py) with the launch schedule of the assignment's code.js -- and
(4) stores inputs + outputs.  The parity tests then check the oracle build (any machine) and the
CUDA library (GPU box) against these files; neither needs /root/reference.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import golden_io as G  # noqa: E402
from oracle import host as OH  # noqa: E402
from oracle import refcl as OR  # noqa: E402

REF = "/root/reference"
A = {1: "Assign01-Sphere_Ray_Tracing", 2: "Assign02-Multi_Sphere_Ray_Tracing", 3: "Assign03-Two_Kernel_Ray_Tracing",
     4: "Assign04-Triangle_Mesh", 5: "Assign05-Bounding_Box", 6: "Assign06-1D_uniform_slab_acceleration",
     7: "Assign07-3D_uniform_grid_acceleration", 8: "Assign08-Shadow_Tracing", 9: "Assign09-Thin_Lens_Camera", 10: "Assign10-Path_Tracing"}


def mesh_arrays(meshes):
    out = {}
    for k, m in enumerate(meshes):
        out["mesh%d_positions" % k] = m["positions"].astype(np.float32)   # fp32-valued already (gl-matrix Float32Array)
        out["mesh%d_normals" % k] = m["normals"].astype(np.float32)
        assert np.array_equal(out["mesh%d_positions" % k].astype(np.float64), m["positions"])
        assert np.array_equal(out["mesh%d_normals" % k].astype(np.float64), m["normals"])
        out["mesh%d_matidx" % k] = m["materialIndices"].astype(np.int32)
        out["mesh%d_materials" % k] = m["materials"]
    return out


def grid_digests(prep_sets):
    out = []
    for s in prep_sets:
        d = {"kind": s["kind"], "n": int(s["n"]), "refs": int(s["box"][-1]), "box": G.digest(s["box"].astype(np.uint32))}
        if s["kind"] == "sphere":
            d["prim"] = G.digest(s["data"])
            d["matid"] = G.digest(s["matid"].astype(np.uint32))
        else:
            d["prim"] = G.digest(s["pos"])
            d["normal"] = G.digest(s["normal"])
            if s["kind"] == "triangle":
                d["matid"] = G.digest(s["matid"].astype(np.uint32))
        out.append(d)
    return out


def gen_a10(lib, scene_file, cols=40, rows=30, rpp=4, passes=2, seed=2015):
    tree, meshes = G.describe_scene(os.path.join(REF, A[10], "scenes", scene_file), OH.parseMeshJSON)
    tmp = tempfile.mkdtemp(prefix="golden_")
    path = G.materialize_scene(tree, meshes, tmp)
    scene = OH.loadScene(path, cols, rows, assignment=10)
    seeds = OR.make_seeds(cols * rows * rpp, seed)
    prep = OR.prepare_a10(scene, 1)
    st = OR.A10State(cols * rows * rpp, seeds)
    lib.a10_initAcu(st.acu, st.total)
    cam16 = scene["camera"].toFloat32Array()
    accum, pixels, seeds_after, counts = [], [], [], []
    for _ in range(passes):
        pix = OR.a10_execute_render(lib, st, prep, cam16, cols, rows, rpp, scene["focal_length"], scene["lens_diameter"])
        acc = np.zeros((cols * rows, 4), np.float32)
        for k in range(rpp):
            acc += st.acu.reshape(cols * rows, rpp, 4)[:, k]
        accum.append(acc)
        pixels.append(pix.copy())
        seeds_after.append(st.seeds.copy())
        counts.append([st.n_closest, st.n_any])
    name = "a10_" + os.path.splitext(scene_file)[0]
    G.save(name, tree=tree, params={"assignment": 10, "cols": cols, "rows": rows, "rpp": rpp, "passes": passes, "seed": seed, "source": scene_file},
           grids=grid_digests(prep["sets"]), cam16=cam16, accum=np.stack(accum), pixels=np.stack(pixels), seeds_after=np.stack(seeds_after),
           counts=np.asarray(counts, dtype=np.int64), **mesh_arrays(meshes))
    print(name, "rays", counts[-1], "mean pixel", float(pixels[-1][..., :3].mean()))


def gen_a089(lib, a, scene_file, cols=64, rows=48, rpp=4):
    tree, _ = G.describe_scene(os.path.join(REF, A[a], "scenes", scene_file), OH.parseMeshJSON)
    path = G.materialize_scene(tree, [], tempfile.mkdtemp(prefix="golden_"))
    scene = OH.loadScene(path, cols, rows, assignment=a)
    if a == 8:
        acu, pix, st = OR.a08_render(lib, scene, cols, rows)
        rpp = 1
    else:
        acu, pix, st = OR.a09_render(lib, scene, cols, rows, rpp)
    name = "a%02d_%s" % (a, os.path.splitext(scene_file)[0])
    hit = st["pois"]["matId"].astype(np.int32)
    G.save(name, tree=tree, params={"assignment": a, "cols": cols, "rows": rows, "rpp": rpp, "n_slabs": 5, "source": scene_file},
           cam16=scene["camera"].toFloat32Array(), acu=acu.astype(np.float32), pixels=pix, matid=hit, maxt=st["rays"]["maxt"].copy())
    print(name, "mean pixel", float(pix[..., :3].mean()), "hits", int((hit >= 0).sum()))


def a456_outputs(lib, out, cols, rows, slabs, mol=None, mesh=None, prefix=""):
    """A04 / A05 (brute force, + boxes) and A06 (x slabs) frames of the same inputs."""
    p, r = OR.a04_render(lib, cols, rows, molData=mol, meshData=mesh)
    out[prefix + "a04_pixels"], out[prefix + "a04_maxt"] = p, r["maxt"].copy()
    p, r = OR.a05_render(lib, cols, rows, molData=mol, meshData=mesh)
    out[prefix + "a05_pixels"], out[prefix + "a05_maxt"] = p, r["maxt"].copy()
    if mol is not None and mesh is None:
        out[prefix + "a04_raytrace"] = OR.a04_raytrace(lib, mol, cols, rows)
    grids = []
    for n in slabs:
        p, r, prep = OR.a06_render(lib, cols, rows, n, molData=mol, meshData=mesh)
        out[prefix + "a06_pixels_n%d" % n], out[prefix + "a06_maxt_n%d" % n] = p, r["maxt"].copy()
        g = {"n": n}
        if "mol" in prep:
            m = prep["mol"]
            g["mol"] = {"refs": int(m["box"][-1]), "box": G.digest(m["box"]), "prim": G.digest(m["atoms"]), "colors": G.digest(m["colors"]),
                        "index": G.digest(m["index"])}
        if "mesh" in prep:
            t = prep["mesh"]
            g["mesh"] = {"refs": int(t["box"][-1]), "box": G.digest(t["box"]), "prim": G.digest(t["pos"]), "normal": G.digest(t["normal"]),
                         "index": G.digest(t["index"])}
        grids.append(g)
    return grids


def load_mol(a, fname):
    with open(os.path.join(REF, A[a], "mol", fname), "r") as f:
        serial, elem, xyz = G.describe_pdb(f.read())
    return serial, elem, xyz, OH.parsePDB(G.pdb_text(serial, elem, xyz))


def gen_mol(lib, fname, cols=64, rows=48, slabs=(2, 5)):
    serial, elem, xyz, mol = load_mol(7, fname)
    out = {"a02_pixels": OR.a02_render(lib, mol, cols, rows)}
    p3, r3 = OR.a03_render(lib, mol, cols, rows)
    out["a03_pixels"], out["a03_mint"] = p3, r3["mint"].copy()
    grids = []
    for n in slabs:
        p7, r7, prep = OR.a07_render(lib, cols, rows, n, molData=mol)
        out["a07_pixels_n%d" % n], out["a07_maxt_n%d" % n] = p7, r7["maxt"].copy()
        m = prep["mol"]
        grids.append({"n": n, "refs": int(m["box"][-1]), "box": G.digest(m["box"]), "prim": G.digest(m["atoms"]), "index": G.digest(m["index"])})
    slab_grids = a456_outputs(lib, out, cols, rows, slabs, mol=mol)
    name = "mol_" + os.path.splitext(fname)[0]
    G.save(name, params={"cols": cols, "rows": rows, "slabs": list(slabs), "size": int(mol["size"]), "source": fname}, grids=grids, slab_grids=slab_grids,
           serial=serial.astype(np.int32), elem=elem.astype("U2"), xyz=xyz, **out)
    print(name, "size", mol["size"], "records", len(serial), "refs", [g["refs"] for g in grids])


def gen_tri(lib, fname, cols=64, rows=48, slabs=(2, 10), grid_slabs=(1, 2, 5, 10, 32), keep_normals=True, with_mol=None):
    jm = OH.parseMeshJSON(os.path.join(REF, A[10] if fname in ("boxes.json", "Cornell_box_model.json") else A[7], "tri", fname))
    mesh = {"positions": np.asarray(jm["positions"], dtype=np.float64), "normals": np.asarray(jm["normals"], dtype=np.float64),
            "materialIndices": np.asarray(jm["materialIndices"], dtype=np.int64), "materials": np.asarray(jm["materials"], dtype=np.float64)}
    if not keep_normals:
        mesh["normals"] = np.zeros_like(mesh["positions"])
    tmp = tempfile.mkdtemp(prefix="golden_")
    p = os.path.join(tmp, "m.json")
    with open(p, "w") as f:
        f.write(G.mesh_json_text(mesh["positions"], mesh["normals"], mesh["materialIndices"], mesh["materials"]))
    md = OH.parseMeshJSON(p)
    out, grids = {}, []
    for n in grid_slabs:
        pos, nor, box, idx = OH.splitMeshData(md, n)
        grids.append({"n": n, "refs": int(box[-1]), "box": G.digest(box.astype(np.uint32)), "prim": G.digest(OH.to_f32(pos)),
                      "normal": G.digest(OH.to_f32(nor)), "index": G.digest(np.asarray(idx, dtype=np.uint32))})
    for n in slabs:
        p7, r7, _ = OR.a07_render(lib, cols, rows, n, meshData=md)
        out["a07_pixels_n%d" % n], out["a07_maxt_n%d" % n] = p7, r7["maxt"].copy()
    params = {"cols": cols, "rows": rows, "slabs": list(slabs), "source": fname, "with_mol": with_mol}
    if with_mol:   # computeBoth (A07/code.js:629-668): molecule then mesh over one ray buffer
        serial, elem, xyz, mol = load_mol(7, with_mol)
        pb, rb, _ = OR.a07_render(lib, cols, rows, 5, molData=mol, meshData=md)
        out.update(both_pixels=pb, both_maxt=rb["maxt"].copy(), both_serial=serial.astype(np.int32), both_elem=elem.astype("U2"), both_xyz=xyz)
    slab_grids = a456_outputs(lib, out, cols, rows, slabs, mesh=md)
    if with_mol:   # computeBoth of A04 / A05 / A06
        a456_outputs(lib, out, cols, rows, (5,), mol=mol, mesh=md, prefix="both_")
    name = "tri_" + os.path.splitext(fname)[0]
    G.save(name, params=params, grids=grids, slab_grids=slab_grids, **mesh_arrays([mesh]), **out)
    print(name, "triangles", md["nTriangles"], "refs", [g["refs"] for g in grids])


def gen_a01(lib):
    out = {}
    for cols, rows in ((64, 64), (50, 30)):
        out["pixels_%dx%d" % (cols, rows)] = OR.a01_render(lib, cols, rows)
    G.save("a01", params={"sizes": [[64, 64], [50, 30]]}, **out)
    print("a01", {k: float(v[..., 0].mean()) for k, v in out.items()})


def main():
    assert os.path.isdir(REF), "run where /root/reference exists"
    lib = OR.load_reference()
    assert lib.kind == "reference"
    gen_a01(lib)
    for f in ("dna.pdb", "benzene.pdb", "c60.pdb"):
        gen_mol(lib, f)
    gen_mol(lib, "3IZ4.pdb", cols=32, rows=24, slabs=(10,))
    gen_tri(lib, "teapot.json", with_mol="dna.pdb")
    gen_tri(lib, "house.json")
    gen_tri(lib, "boxes.json", slabs=(5,))
    gen_tri(lib, "house_of_parliament.json", cols=48, rows=36, slabs=(10,), keep_normals=False)
    for a in (8, 9):
        for f in sorted(os.listdir(os.path.join(REF, A[a], "scenes"))):
            gen_a089(lib, a, f)
    for f in sorted(os.listdir(os.path.join(REF, A[10], "scenes"))):
        gen_a10(lib, f)


if __name__ == "__main__":
    main()
