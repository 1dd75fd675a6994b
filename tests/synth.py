"""Synthetic, self-made test assets in the reference's on-disk formats (no reference data):
a displaced-sphere triangle mesh in the ``tri/*.json`` schema, Cornell-style XML scenes in the
``scenes/*.xml`` schema (with BOM and commented-out geometry, like the reference's files),
and a PDB-format molecule.  Shared by tests/, bench.py and __graft_entry__.smoke()."""
from __future__ import annotations

import os

import numpy as np


def synth_mesh(n_u=40, n_v=20, seed=2015, model_matrix=None):
    """Unit sphere displaced radially by a small sin-sum height field, n_u x n_v quads =
    2*n_u*n_v triangles, smooth per-vertex normals; one node, one material.  Returned in the
    tri/*.json object model (numpy arrays where JSON has number lists)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ph = rng.uniform(0, 2 * np.pi, size=6)
    u = np.linspace(0.0, 2 * np.pi, n_u + 1)
    v = np.linspace(0.02, np.pi - 0.02, n_v + 1)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    h = 1.0 + 0.05 * (np.sin(3 * uu + ph[0]) * np.sin(4 * vv + ph[1]) + 0.5 * np.sin(7 * uu + ph[2]) * np.sin(5 * vv + ph[3])
                      + 0.25 * np.sin(13 * uu + ph[4]) * np.sin(11 * vv + ph[5]))
    x = h * np.sin(vv) * np.cos(uu)
    y = h * np.cos(vv)
    z = h * np.sin(vv) * np.sin(uu)
    P = np.stack([x, y, z], axis=-1)
    # smooth normals from central differences of the parametrisation
    du = np.roll(P, -1, axis=1) - np.roll(P, 1, axis=1)
    du[:, 0] = P[:, 1] - P[:, -2]
    du[:, -1] = P[:, 1] - P[:, -2]
    dv = np.empty_like(P)
    dv[1:-1] = P[2:] - P[:-2]
    dv[0] = P[1] - P[0]
    dv[-1] = P[-1] - P[-2]
    N = np.cross(du, dv)
    N /= np.maximum(np.linalg.norm(N, axis=-1, keepdims=True), 1e-12)
    N *= np.sign(np.sum(N * P, axis=-1, keepdims=True) + 1e-30)
    W = n_u + 1
    i, j = np.meshgrid(np.arange(n_v), np.arange(n_u), indexing="ij")
    a = (i * W + j).ravel()
    b = (i * W + j + 1).ravel()
    c = ((i + 1) * W + j).ravel()
    d = ((i + 1) * W + j + 1).ravel()
    # counter-clockwise seen from outside (the reference's interTriangle is one-sided)
    idx = np.stack([a, b, c, b, d, c], axis=1).reshape(-1)
    mm = model_matrix if model_matrix is not None else [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    # fp32-representable inputs keep the JSON text short and the file exactly re-loadable
    return {
        "name": "synth",
        "materials": [{"diffuseReflectance": [0.8, 0.7, 0.5, 1]}],
        "meshes": [{"vertexPositions": P.reshape(-1).astype(np.float32).astype(np.float64),
                    "vertexNormals": N.reshape(-1).astype(np.float32).astype(np.float64),
                    "indices": idx.astype(np.int64), "materialIndex": 0}],
        "nodes": [{"modelMatrix": list(mm), "meshIndices": [0]}],
    }


def mesh_to_json_file(mesh, path):
    import json
    m = dict(mesh)
    m["meshes"] = [{k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in mm.items()} for mm in mesh["meshes"]]
    with open(path, "w") as f:
        json.dump(m, f)


_XML_HEAD = """﻿<?xml version="1.0" encoding="UTF-8"?>
<scene>
<!-- synthetic scene in the scenes/*.xml schema of the reference (written for this repo) -->
<camera>
	<eye><x>0.0</x><y>0.1</y><z>2.2</z></eye>
	<lookAt><x>0.0</x><y>-0.15</y><z>0.0</z></lookAt>
	<vup><x>0.0</x><y>1.0</y><z>0.0</z></vup>
	<fov>55</fov>
	<focal_length>2.1</focal_length>
	<lens_diameter>0.02</lens_diameter>
</camera>
"""


def _v(tag, x, y, z):
    return "<%s><x>%r</x><y>%r</y><z>%r</z></%s>" % (tag, x, y, z, tag)


def _light(pos, nor, irr, rad):
    return "<light>%s%s%s<radius>%r</radius></light>\n" % (_v("position", *pos), _v("normal", *nor), _v("irradiance", *irr), rad)


def _mat(name, r, g, b):
    return "<material><id>%s</id><color><r>%r</r><g>%r</g><b>%r</b><a>1.0</a></color></material>\n" % (name, r, g, b)


def _tri(p0, p1, p2, n, mat):
    return "<triangle>%s%s%s%s%s%s<matId>%s</matId></triangle>\n" % (
        _v("p0", *p0), _v("p1", *p1), _v("p2", *p2), _v("n0", *n), _v("n1", *n), _v("n2", *n), mat)


def _quad(a, b, c, d, n, mat):
    return _tri(a, b, c, n, mat) + _tri(a, c, d, n, mat)


def synth_scene_xml(n_lights=2, with_sphere=True, with_mesh=True, mesh_nslabs=12, mesh_file="./tri/synth.json", closed=True):
    """A Cornell-style box: five (or six) walls of two one-sided triangles each facing inward,
    1-2 disk lights, optionally a sphere and a <mesh>.  Windings follow the reference's
    one-sided interTriangle (dot(cross(e2,e1), d) > 0 is a front hit)."""
    s = _XML_HEAD
    s += _light((0.0, 0.8, 0.1), (0.0, -1.0, 0.0), (25, 25, 25), 0.12)
    if n_lights > 1:
        s += _light((0.65, 0.25, 0.55), (-1.0, -0.8, -1.0), (20, 22, 25), 0.1)   # un-normalised normal, like cornell_teapot3
    if n_lights > 2:
        s += _light((-0.6, 0.5, 0.6), (1.0, -1.0, -0.5), (15, 10, 10), 0.08)
    s += _mat("red", 0.85, 0.2, 0.2) + _mat("green", 0.2, 0.85, 0.25) + _mat("white", 0.8, 0.8, 0.8) + _mat("blue", 0.15, 0.2, 0.9)
    s += _mat("cream", 0.89, 0.85, 0.79)
    s += "<!--\n<sphere><center><x>9</x><y>9</y><z>9</z></center><radius>1</radius><matId>red</matId></sphere>\n-->\n"
    if with_sphere:
        s += "<sphere>%s<radius>0.22</radius><matId>blue</matId></sphere>\n" % _v("center", -0.45, -0.76, 0.25)
    # Walls sit at +-A but SPAN +-B (> A), so every wall lies strictly inside the triangle-set AABB:
    # the reference accepts a hit only with t < the ray's AABB exit parameter, and a wall lying exactly
    # on a bounds face is hit or missed by rounding (its own scenes keep walls at 0.99 inside a +-1 box).
    A, B, ZF, ZB = 0.98, 1.0, 2.5, 2.6
    # floor (normal +y), ceiling (-y), back (+z), left (+x), right (-x), front (-z, behind the camera, optional)
    walls = [
        ((-B, -A, ZB), (B, -A, ZB), (B, -A, -B), (-B, -A, -B), (0.0, 1.0, 0.0), "white"),
        ((-B, A, -B), (B, A, -B), (B, A, ZB), (-B, A, ZB), (0.0, -1.0, 0.0), "white"),
        ((-B, -B, -A), (B, -B, -A), (B, B, -A), (-B, B, -A), (0.0, 0.0, 1.0), "white"),
        ((-A, -B, ZB), (-A, -B, -B), (-A, B, -B), (-A, B, ZB), (1.0, 0.0, 0.0), "red"),
        ((A, -B, -B), (A, -B, ZB), (A, B, ZB), (A, B, -B), (-1.0, 0.0, 0.0), "green"),
    ]
    if closed:
        walls.append(((B, -B, ZF), (-B, -B, ZF), (-B, B, ZF), (B, B, ZF), (0.0, 0.0, -1.0), "white"))
    for a, b, c, d, n, m in walls:
        s += _quad(a, b, c, d, n, m)
    if with_mesh:
        s += ("<mesh><file>%s</file><nslabs>%d</nslabs><normalize>yes</normalize>%s%s<matId>cream</matId></mesh>\n"
              % (mesh_file, mesh_nslabs, _v("scale", 0.7, 0.7, 0.7), _v("translate", 0.3, -0.5, 0.1)))
    s += "</scene>\n"
    return s


def write_scene(tmpdir, **kw):
    """Writes <tmpdir>/scenes/synth.xml (+ returns its path); the mesh is supplied through the
    loaders' mesh_loader hook, so no JSON needs to be written for large meshes."""
    d = os.path.join(str(tmpdir), "scenes")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, "synth.xml")
    with open(p, "w", encoding="utf-8") as f:
        f.write(synth_scene_xml(**kw))
    return p


def synth_pdb(n_atoms=200, seed=2015, gap_at=57):
    """PDB text with ATOM/HETATM records, one serial-number gap (a TER record consumes a
    serial, as in the reference's 3IZ4.pdb) so that parsePDB's size exceeds its record count."""
    rng = np.random.Generator(np.random.PCG64(seed))
    elems = ["C", "N", "O", "H", "P", "S"]
    lines = ["HEADER    SYNTHETIC MOLECULE"]
    serial = 1
    for i in range(n_atoms):
        if i == gap_at:
            lines.append("TER   %5d" % serial)
            serial += 1
        e = elems[int(rng.integers(0, len(elems)))]
        x, y, z = rng.uniform(-12, 12, size=3)
        rec = "ATOM  " if i % 7 else "HETATM"
        lines.append("%s%5d %-4s RES A%4d    %8.3f%8.3f%8.3f  1.00  0.00          %2s" % (rec, serial, e, i % 999, x, y, z, e))
        serial += 1
    lines.append("END")
    return "\n".join(lines) + "\n"
