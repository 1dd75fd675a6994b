"""Golden-fixture plumbing shared by tests/golden/make_golden.py (the generator, run where
/root/reference exists) and the parity tests (run anywhere, GPU box included).

A fixture is ONE ``.npz`` under tests/golden/ holding
  * the INPUT of a reference demo in neutral numeric form -- never a copy of a reference file:
      - XML scenes as a JSON element tree of tags and number strings (comments/BOM dropped),
      - meshes as the float32 triangle soup ``parseMeshJSON`` produces (positions / normals after
        the node transform, which gl-matrix already rounded to fp32),
      - molecules as (serial, element, x, y, z) records,
  * the canvas size / sampling parameters / seed, and
  * the OUTPUT of the oracle (the reference's own kernels + the JS-order host restatement) on
    that input: accumulation image, seed buffer after the pass, pixel image, ray counts, grid
    cell-list digests.

``materialize_*`` re-emits the input as files in the reference's on-disk formats (in a temp
directory), so both the oracle's and the product's own loaders are exercised on it.
"""
from __future__ import annotations

import hashlib
import json
import os
import xml.etree.ElementTree as ET

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ------------------------------------------------------------------------------ XML scenes
def xml_tree(elem):
    kids = list(elem)
    if not kids:
        return [elem.tag, (elem.text or "").strip()]
    return [elem.tag, [xml_tree(k) for k in kids]]


def tree_xml(t, depth=0):
    tag, body = t
    if isinstance(body, str):
        return "%s<%s>%s</%s>\n" % (" " * depth, tag, body, tag)
    return "%s<%s>\n%s%s</%s>\n" % (" " * depth, tag, "".join(tree_xml(k, depth + 1) for k in body), " " * depth, tag)


def describe_scene(xml_path, parse_mesh):
    """(tree, meshes): the scene's element tree with every <mesh><file> replaced by a neutral name,
    and the parsed meshes (``parse_mesh(path)`` = the oracle's parseMeshJSON)."""
    with open(xml_path, "r", encoding="utf-8-sig") as f:
        root = ET.fromstring(f.read())
    base = os.path.dirname(os.path.dirname(os.path.abspath(xml_path)))
    meshes = []
    for xm in root.iter("mesh"):
        f = xm.find("file")
        jm = parse_mesh(os.path.join(base, f.text))
        k = len(meshes)
        meshes.append({"positions": np.asarray(jm["positions"], dtype=np.float64), "normals": np.asarray(jm["normals"], dtype=np.float64),
                       "materialIndices": np.asarray(jm["materialIndices"], dtype=np.int64),
                       "materials": np.asarray(jm["materials"], dtype=np.float64)})
        f.text = "./tri/golden_mesh%d.json" % k
    return xml_tree(root), meshes


def mesh_json_text(positions, normals, material_indices=None, materials=None):
    """A tri/*.json model (assimp-style schema of the reference's loader) holding a triangle soup:
    one mesh per material index, non-indexed vertices, one identity node.  Values are written with
    repr() so fp32-valued inputs reload exactly."""
    p = np.asarray(positions, dtype=np.float64).reshape(-1, 9)
    n = np.asarray(normals, dtype=np.float64).reshape(-1, 9)
    mi = np.zeros(len(p), dtype=np.int64) if material_indices is None else np.asarray(material_indices, dtype=np.int64)
    mats = np.asarray([0.8, 0.8, 0.8, 1.0] if materials is None else materials, dtype=np.float64).reshape(-1, 4)
    # keep the triangle order: consecutive runs of one material become one mesh each
    meshes, run_start = [], 0
    for i in range(1, len(p) + 1):
        if i == len(p) or mi[i] != mi[run_start]:
            meshes.append({"vertexPositions": [float(v) for v in p[run_start:i].reshape(-1)],
                           "vertexNormals": [float(v) for v in n[run_start:i].reshape(-1)], "materialIndex": int(mi[run_start])})
            run_start = i
    model = {"materials": [{"diffuseReflectance": [float(v) for v in m]} for m in mats], "meshes": meshes,
             "nodes": [{"modelMatrix": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1], "meshIndices": list(range(len(meshes)))}]}
    return json.dumps(model)


def materialize_scene(tree, meshes, tmpdir):
    """Writes <tmpdir>/scenes/golden.xml (with a BOM and a commented-out element, like the
    reference's files) and <tmpdir>/tri/golden_mesh<k>.json; returns the XML path."""
    tmpdir = str(tmpdir)
    os.makedirs(os.path.join(tmpdir, "scenes"), exist_ok=True)
    os.makedirs(os.path.join(tmpdir, "tri"), exist_ok=True)
    body = tree_xml(tree)
    body = body.replace("<scene>\n", "<scene>\n<!--\n<sphere><center><x>9</x><y>9</y><z>9</z></center><radius>1</radius></sphere>\n-->\n", 1)
    path = os.path.join(tmpdir, "scenes", "golden.xml")
    with open(path, "w", encoding="utf-8-sig") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n' + body)
    for k, m in enumerate(meshes):
        with open(os.path.join(tmpdir, "tri", "golden_mesh%d.json" % k), "w", encoding="utf-8") as f:
            f.write(mesh_json_text(m["positions"], m["normals"], m.get("materialIndices"), m.get("materials")))
    return path


# ------------------------------------------------------------------------------ molecules
def describe_pdb(text):
    """(serial, element, x, y, z) of every record the reference's parser accepts
    (mol/pdbParserV1.js:20-36: ATOM/HETATM, altLoc ' ' or 'A')."""
    serial, elem, xyz = [], [], []
    for raw in text.split("\n"):
        line = raw.lstrip()
        if line[0:6] not in ("ATOM  ", "HETATM") or line[16:17] not in (" ", "A"):
            continue
        serial.append(int(line[6:11]))
        elem.append(line[76:78].replace(" ", "") or line[12:16].replace(" ", ""))
        xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.asarray(serial, dtype=np.int64), np.asarray(elem), np.asarray(xyz, dtype=np.float64).reshape(-1, 3)


def pdb_text(serial, elem, xyz):
    """Re-emits the records in PDB column layout (coordinates are %8.3f in the format itself)."""
    lines = ["HEADER    GOLDEN FIXTURE"]
    for s, e, (x, y, z) in zip(serial, elem, xyz):
        lines.append("ATOM  %5d %-4s RES A   1    %8.3f%8.3f%8.3f  1.00  0.00          %2s" % (int(s), str(e), x, y, z, str(e)))
    lines.append("END")
    return "\n".join(lines) + "\n"


# ------------------------------------------------------------------------------ storage
def digest(a) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.view(np.uint8).reshape(-1).tobytes()).hexdigest()


def save(name, **arrays):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    out = {}
    for k, v in arrays.items():
        if isinstance(v, (dict, list)) and not isinstance(v, np.ndarray):
            out[k + "__json"] = np.frombuffer(json.dumps(v).encode("utf-8"), dtype=np.uint8)
        else:
            out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        if k.endswith("__json"):
            out[k[:-6]] = json.loads(bytes(z[k]).decode("utf-8"))
        else:
            out[k] = z[k]
    return out


def names(prefix=""):
    if not os.path.isdir(GOLDEN_DIR):
        return []
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix))


def meshes_of(fx):
    out, k = [], 0
    while "mesh%d_positions" % k in fx:
        out.append({"positions": fx["mesh%d_positions" % k].astype(np.float64), "normals": fx["mesh%d_normals" % k].astype(np.float64),
                    "materialIndices": fx["mesh%d_matidx" % k], "materials": fx["mesh%d_materials" % k]})
        k += 1
    return out
