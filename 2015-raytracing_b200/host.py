"""Host-side mirror of the reference's ``code.js`` surface on top of the C ABI.

Same names, argument meaning and error behaviour as the JavaScript host of
eaymerich/2015-RayTracing (citations relative to /root/reference, A10 =
Assign10-Path_Tracing): the ``mol/`` ``tri/`` ``scenes/`` loaders, ``Bounds``, ``Camera``,
``Light``, ``bounds2AABB``, the ``split*Data`` grid builders (here: integer CUDA kernels
through ``rt_grid_build_*``), ``Mesh`` and the ``preRender / executeRender / postRender``
frame driver.  JS ``Number`` is an IEEE double = Python ``float``; ``Float32Array`` stores are
``numpy.float32`` casts at the same points as in the reference.

Python stands in for the headless Node.js layer because this image has no Node toolchain
(INTEGRATION.md shows the N-API addon that binds the same C entry points).
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import re
import xml.etree.ElementTree as ET

import numpy as np

from . import lib as L

_MAX = 1.7976931348623157e308  # Number.MAX_VALUE


# ------------------------------------------------------------------------------ small types
class Bounds:
    """``Bounds`` of lib/utilities.js (A10/lib/utilities.js:389-422)."""

    def __init__(self, min=None, max=None):
        self.min = [_MAX, _MAX, _MAX] if min is None else [float(min[0]), float(min[1]), float(min[2])]
        self.max = [-_MAX, -_MAX, -_MAX] if max is None else [float(max[0]), float(max[1]), float(max[2])]

    def center(self):
        return [(a + b) / 2 for a, b in zip(self.min, self.max)]

    def diagonal(self):
        d = [b - a for a, b in zip(self.min, self.max)]
        return math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])

    def merge(self, b):
        self.min = [min(x, y) for x, y in zip(self.min, b.min)]
        self.max = [max(x, y) for x, y in zip(self.max, b.max)]


def bounds2AABB(bounds) -> np.ndarray:
    """A10/code.js:610-621."""
    with np.errstate(over="ignore"):
        return np.array([*bounds.min, 1.0, *bounds.max, 1.0], dtype=np.float64).astype(np.float32)


class Vec3:
    """A10/code.js:13-53."""

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def set(self, x, y, z):
        self.x, self.y, self.z = x, y, z

    def subtract(self, b):
        return Vec3(self.x - b.x, self.y - b.y, self.z - b.z)

    def cross(self, b):
        return Vec3(self.y * b.z - self.z * b.y, self.z * b.x - self.x * b.z, self.x * b.y - self.y * b.x)

    def normalize(self):
        n = math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)
        with np.errstate(all="ignore"):
            self.x, self.y, self.z = (float(np.float64(c) / np.float64(n)) for c in (self.x, self.y, self.z))

    def tolist(self):
        return [self.x, self.y, self.z]


class Camera:
    """A10/code.js:175-277."""

    def __init__(self):
        self.eye, self.U, self.V, self.W = Vec3(), Vec3(), Vec3(), Vec3()
        self.width = self.height = 1.0
        self.cols = self.rows = 0

    def defaultInit(self):
        self.eye, self.U, self.V, self.W = Vec3(0, 0, 0), Vec3(1, 0, 0), Vec3(0, 1, 0), Vec3(0, 0, 1)

    def _frustum(self, fov, cols, rows):
        self.cols, self.rows = cols, rows
        self.height = 2.0 * math.tan(0.5 * fov * math.pi / 180.0)
        self.width = self.height * (cols / rows)

    def set(self, bounds, cols, rows):
        self._frustum(60, cols, rows)
        c = bounds.center()
        self.eye.set(c[0], c[1], c[2] + bounds.diagonal())

    def lookAt(self, eye, lookat, vup, fov, cols, rows):
        self._frustum(fov, cols, rows)
        self.eye = eye
        self.W = eye.subtract(lookat)
        self.W.normalize()
        self.U = vup.cross(self.W)
        self.U.normalize()
        self.V = self.W.cross(self.U)

    def rotate(self, bounds, angle):
        c, diag = bounds.center(), bounds.diagonal()
        rad = angle * math.pi / 180.0
        self.eye.set(c[0] + math.sin(rad) * diag, c[1], c[2] + math.cos(rad) * diag)
        self.W.set(self.eye.x - c[0], self.eye.y - c[1], self.eye.z - c[2])
        self.W.normalize()
        self.U = self.V.cross(self.W)

    def toFloat32Array(self) -> np.ndarray:
        return np.array([*self.eye.tolist(), *self.U.tolist(), *self.V.tolist(), *self.W.tolist(), self.width, self.height,
                         self.cols, self.rows], dtype=np.float64).astype(np.float32)


class Light:
    """Disk area light, A10/code.js:279-353."""

    def __init__(self):
        self.position, self.normal, self.T, self.B, self.irradiance = Vec3(), Vec3(), Vec3(), Vec3(), Vec3()
        self.radius = 0.0
        self.area = 0.0

    def set(self, position, normal, irradiance, radius):
        self.position, self.normal, self.irradiance, self.radius = position, normal, irradiance, radius
        self.normal.normalize()
        self.calculateArea()
        self.calculateTBN()

    def calculateArea(self):
        self.area = math.pi * self.radius * self.radius

    def calculateTBN(self):
        n = self.normal
        mags = [abs(n.x), abs(n.y), abs(n.z)]
        v = Vec3(n.x, n.y, n.z)
        m = min(mags)
        if m == mags[0]:
            v.x = 1.0
        elif m == mags[1]:
            v.y = 1.0
        else:
            v.z = 1.0
        v.normalize()
        self.T = v.cross(n)
        self.T.normalize()
        self.B = n.cross(self.T)
        self.B.normalize()

    def _info(self, b, c, s):
        return np.array([*self.position.tolist(), *b.tolist(), *c.tolist(), s, 0, 0, 0, 0, 0, 0], dtype=np.float64).astype(np.float32)

    def toShadowInfo(self):
        return self._info(self.T, self.B, self.radius)

    def toSceneRenderInfo(self):
        return self._info(self.normal, self.irradiance, self.area)

    def toLightRenderInfo(self):
        return self._info(self.normal, self.irradiance, self.radius)


# ------------------------------------------------------------------------------ loaders
def parseMeshJSON(jsonFileName):
    """tri/meshDataVersion1.js (A10/tri/meshDataVersion1.js:12-78), vectorised.  gl-matrix
    2.2.1 stores matrices and transformed vectors in Float32Array
    (A10/lib/gl-matrix.js:79-80), so every vertex/normal is rounded to fp32 here, before the
    grid build sees it; the arithmetic itself is float64, left to right."""
    if isinstance(jsonFileName, dict):
        model = jsonFileName
    else:
        with open(jsonFileName, "r", encoding="utf-8-sig") as f:
            model = json.load(f)
    nodes = model.get("nodes")
    pos_parts, nor_parts, mat_parts = [], [], []
    b = Bounds()
    for k in range(len(nodes) if nodes else 1):
        if nodes:
            m = np.asarray(nodes[k]["modelMatrix"], dtype=np.float64).astype(np.float32).astype(np.float64)
            mesh_ids = nodes[k]["meshIndices"]
        else:
            m = np.eye(4, dtype=np.float64).reshape(-1)
            mesh_ids = range(len(model["meshes"]))
        nm = _normalFromMat4(m)
        for index in mesh_ids:
            mesh = model["meshes"][index]
            vp = np.asarray(mesh["vertexPositions"], dtype=np.float64).reshape(-1, 3)
            vn = np.asarray(mesh["vertexNormals"], dtype=np.float64).reshape(-1, 3)
            tp = _xform4(vp, m)
            if len(tp):
                lo, hi = tp.min(axis=0), tp.max(axis=0)
                for a in range(3):
                    if lo[a] < b.min[a]:
                        b.min[a] = float(lo[a])
                    if hi[a] > b.max[a]:
                        b.max[a] = float(hi[a])
            ind = mesh.get("indices")
            idx = np.asarray(ind, dtype=np.int64) if (ind is not None and len(ind) > 0) else np.arange(len(vp), dtype=np.int64)
            nT = len(idx) // 3
            idx = idx[: nT * 3]
            pos_parts.append(tp[idx].reshape(nT, 9))
            nor_parts.append(_xform3(vn[idx], nm).reshape(nT, 9))
            mat_parts.append(np.full(nT, mesh["materialIndex"], dtype=np.uint32))
    positions = np.concatenate(pos_parts) if pos_parts else np.zeros((0, 9))
    normals = np.concatenate(nor_parts) if nor_parts else np.zeros((0, 9))
    matidx = np.concatenate(mat_parts) if mat_parts else np.zeros(0, np.uint32)
    materials = [c for mt in model["materials"] for c in mt["diffuseReflectance"][:4]]
    return {"nTriangles": int(len(positions)), "nMaterials": len(model["materials"]), "materialIndices": matidx,
            "materials": materials, "bounds": b, "positions": np.ascontiguousarray(positions),
            "normals": np.ascontiguousarray(normals), "tCoords": None}


def parseMeshJSON_native(source):
    """parseMeshJSON through the library's native parser (rt_parse_mesh_json): `source` is a file name or the JSON
    text as bytes.  Same result object as :func:`parseMeshJSON`, ~an order of magnitude faster on big meshes."""
    if isinstance(source, (bytes, bytearray)):
        text = bytes(source)
    else:
        with open(source, "rb") as f:
            text = f.read()
    md = L.MeshData()
    rc = L.dll.rt_parse_mesh_json(text, len(text), C.byref(md))
    if rc != 0:
        raise ValueError("parseMeshJSON: malformed model (rt_parse_mesh_json returned %d)" % rc)
    try:
        nt, nm = int(md.n_triangles), int(md.n_materials)
        pos = np.ctypeslib.as_array(md.positions, shape=(max(nt, 1) * 9,))[:nt * 9].reshape(nt, 9).copy()
        nor = np.ctypeslib.as_array(md.normals, shape=(max(nt, 1) * 9,))[:nt * 9].reshape(nt, 9).copy()
        idx = np.ctypeslib.as_array(md.material_indices, shape=(max(nt, 1),))[:nt].astype(np.uint32)
        mats = [float(v) for v in np.ctypeslib.as_array(md.materials, shape=(max(nm, 1) * 4,))[:nm * 4]]
        b = Bounds(list(md.bounds_min), list(md.bounds_max))
    finally:
        L.dll.rt_mesh_data_free(C.byref(md))
    return {"nTriangles": nt, "nMaterials": nm, "materialIndices": idx, "materials": mats, "bounds": b, "positions": pos, "normals": nor,
            "tCoords": None}


def parsePDB_native(text):
    """parsePDB through the library's native parser (rt_parse_pdb); same result object as :func:`parsePDB`."""
    raw = text.encode("latin-1") if isinstance(text, str) else bytes(text)
    md = L.MolData()
    rc = L.dll.rt_parse_pdb(raw, len(raw), C.byref(md))
    if rc != 0:
        raise ValueError("parsePDB: unsupported record (rt_parse_pdb returned %d)" % rc)
    try:
        nr, ne = int(md.n_records), int(md.n_elements)
        ad = np.ctypeslib.as_array(md.atom_data, shape=(max(nr, 1) * 4,))[:nr * 4].copy()
        atomData = []
        for i in range(nr):
            atomData += [int(ad[4 * i]), float(ad[4 * i + 1]), float(ad[4 * i + 2]), float(ad[4 * i + 3])]
        colorData = [float(v) for v in np.ctypeslib.as_array(md.color_data, shape=(max(ne, 1) * 4,))[:ne * 4]]
        colorData = [int(v) if i % 4 == 3 else v for i, v in enumerate(colorData)]
        radiusData = [float(v) for v in np.ctypeslib.as_array(md.radius_data, shape=(max(ne, 1),))[:ne]]
        out = {"size": int(md.size), "atomData": atomData, "colorData": colorData, "radiusData": radiusData,
               "bounds": Bounds(list(md.bounds_min), list(md.bounds_max))}
    finally:
        L.dll.rt_mol_data_free(C.byref(md))
    return out


def _f32round(a):
    return a.astype(np.float32).astype(np.float64)


def _xform4(v, m):  # vec3.transformMat4, A10/lib/gl-matrix.js:1063-1070
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    return _f32round(np.stack([m[0] * x + m[4] * y + m[8] * z + m[12], m[1] * x + m[5] * y + m[9] * z + m[13],
                               m[2] * x + m[6] * y + m[10] * z + m[14]], axis=1))


def _xform3(v, m):  # vec3.transformMat3, A10/lib/gl-matrix.js:1079-1085
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    return _f32round(np.stack([x * m[0] + y * m[3] + z * m[6], x * m[1] + y * m[4] + z * m[7], x * m[2] + y * m[5] + z * m[8]], axis=1))


def _normalFromMat4(a):  # mat3.normalFromMat4, A10/lib/gl-matrix.js:2723-2760
    a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33 = (float(v) for v in a)
    b00, b01, b02 = a00 * a11 - a01 * a10, a00 * a12 - a02 * a10, a00 * a13 - a03 * a10
    b03, b04, b05 = a01 * a12 - a02 * a11, a01 * a13 - a03 * a11, a02 * a13 - a03 * a12
    b06, b07, b08 = a20 * a31 - a21 * a30, a20 * a32 - a22 * a30, a20 * a33 - a23 * a30
    b09, b10, b11 = a21 * a32 - a22 * a31, a21 * a33 - a23 * a31, a22 * a33 - a23 * a32
    det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06
    if not det:
        raise ValueError("parseMeshJSON: singular modelMatrix (the reference throws on the null normal matrix)")
    det = 1.0 / det
    out = [(a11 * b11 - a12 * b10 + a13 * b09) * det, (a12 * b08 - a10 * b11 - a13 * b07) * det, (a10 * b10 - a11 * b08 + a13 * b06) * det,
           (a02 * b10 - a01 * b11 - a03 * b09) * det, (a00 * b11 - a02 * b08 + a03 * b07) * det, (a01 * b08 - a00 * b10 - a03 * b06) * det,
           (a31 * b05 - a32 * b04 + a33 * b03) * det, (a32 * b02 - a30 * b05 - a33 * b01) * det, (a30 * b04 - a31 * b02 + a33 * b00) * det]
    return _f32round(np.asarray(out, dtype=np.float64))


_COLORS = {"H": 0xCCCCCC, "C": 0xAAAAAA, "O": 0xCC0000, "N": 0x0000CC, "S": 0xCCCC00, "P": 0x6622CC, "F": 0x00CC00, "CL": 0x00CC00,
           "BR": 0x882200, "I": 0x6600AA, "FE": 0xCC6600, "CA": 0x8888AA}
_RADII = {"H": 1.2, "Li": 1.82, "Na": 2.27, "K": 2.75, "C": 1.7, "N": 1.55, "O": 1.52, "F": 1.47, "P": 1.80, "S": 1.80, "CL": 1.75,
          "BR": 1.85, "SE": 1.90, "ZN": 1.39, "CU": 1.4, "NI": 1.63}
_FLOAT_PREFIX = re.compile(r"\s*[+-]?(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)")


def _parseFloat(s):
    m = _FLOAT_PREFIX.match(s)
    return float(m.group(0)) if m else float("nan")


def parsePDB(text):
    """mol/pdbParserV1.js (A10/mol/pdbParserV1.js:2-85).  ``size`` is the LENGTH of the sparse
    ``atoms[serial-1]`` array (largest serial), ``atomData`` holds (elemIdx,x,y,z) per
    existing atom -- for 3IZ4.pdb size is one more than the record count (quirk Q13)."""
    atoms = {}
    size = 0
    for raw in text.split("\n"):
        line = raw.lstrip()
        rec = line[0:6]
        if rec not in ("ATOM  ", "HETATM"):
            continue
        if line[16:17] not in (" ", "A"):
            continue
        serial = int(line[6:11])
        elem = line[76:78].replace(" ", "") or line[12:16].replace(" ", "")
        atoms[serial - 1] = (elem, _parseFloat(line[30:38]), _parseFloat(line[38:46]), _parseFloat(line[46:54]))
        size = max(size, serial)
    colorData, radiusData, atomData, used = [], [], [], {}
    lo, hi = [_MAX] * 3, [-_MAX] * 3
    for i in sorted(atoms):
        elem, x, y, z = atoms[i]
        if elem not in used:
            h = _COLORS[elem]
            colorData += [((h >> 16) & 255) / 255, ((h >> 8) & 255) / 255, (h & 255) / 255, 1]
            radiusData.append(_RADII[elem])
            used[elem] = len(used)
        R = radiusData[used[elem]]
        atomData += [used[elem], x, y, z]
        for a, v in enumerate((x, y, z)):
            if v - R < lo[a]:
                lo[a] = v - R
            if v + R > hi[a]:
                hi[a] = v + R
    return {"size": size, "atomData": atomData, "colorData": colorData, "radiusData": radiusData, "bounds": Bounds(lo, hi)}


# ------------------------------------------------------------------------------ grid build (GPU)
def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class DeviceGrid:
    """Device-resident output of a split*Data call (cell-ordered primitives + box_size)."""

    def __init__(self, ctx, grid, bounds_aabb, cells=None):
        self.ctx, self.grid, self.aabb = ctx, grid, np.ascontiguousarray(bounds_aabb, dtype=np.float32)
        self.cells = cells   # None: n^3 cells; the 1-D slabs of Assignment 6 have n

    @property
    def n_refs(self):
        return int(self.grid.n_refs)

    @property
    def n_slabs(self):
        return int(self.grid.n_slabs)

    def box_size(self) -> np.ndarray:
        n = self.n_slabs
        return self.ctx.download(self.grid.box_size, np.uint32, (n * n * n if self.cells is None else self.cells) + 1)

    def prim(self) -> np.ndarray:
        per = 4 if self.grid.kind == 0 else 12
        return self.ctx.download(self.grid.prim, np.float32, self.n_refs * per)

    def normal(self) -> np.ndarray:
        return self.ctx.download(self.grid.normal, np.float32, self.n_refs * 12)

    def matid(self) -> np.ndarray:
        return self.ctx.download(self.grid.matid, np.uint32, self.n_refs)

    def release(self):
        if self.grid is not None:
            L.dll.rt_grid_release(self.ctx.h, C.byref(self.grid))
            self.grid = None


def _build_triangles(ctx, pos9, nor9, ids, bounds, n_slabs, xform=None):
    pos9 = np.ascontiguousarray(pos9, dtype=np.float64).reshape(-1, 9)
    nor9 = np.ascontiguousarray(nor9, dtype=np.float64).reshape(-1, 9)
    ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)
    g = L.Grid()
    xf = None
    if xform is not None:
        xf = L.MeshXform(int(xform["normalize"]), _d3(xform["center"]), float(xform["maxdim"]), _d3(xform["scale"]), _d3(xform["translate"]))
    ctx.check(L.dll.rt_grid_build_triangles(ctx.h, L.hptr(pos9), L.hptr(nor9), L.hptr(ids), len(pos9), _d3(bounds.min), _d3(bounds.max),
                                            int(n_slabs), C.byref(xf) if xf is not None else None, C.byref(g)))
    return g


def splitMeshData(ctx, meshData, nn_slabs, xform=None) -> "L.Grid":
    """splitMeshData (A10/code.js:899-1041; A07/code.js:980-1122 adds the per-reference
    material index, returned in ``grid.matid``)."""
    return _build_triangles(ctx, meshData["positions"], meshData["normals"], meshData["materialIndices"], meshData["bounds"], nn_slabs, xform)


def splitTriangleData(ctx, scene, n_slabs) -> "L.Grid":
    """splitTriangleData (A10/code.js:1643-1772)."""
    t = scene["triangles"]
    pos9 = np.array([[*x["p0"].tolist(), *x["p1"].tolist(), *x["p2"].tolist()] for x in t], dtype=np.float64).reshape(-1, 9)
    nor9 = np.array([[*x["n0"].tolist(), *x["n1"].tolist(), *x["n2"].tolist()] for x in t], dtype=np.float64).reshape(-1, 9)
    ids = np.array([x["matId"] for x in t], dtype=np.uint32)
    return _build_triangles(ctx, pos9, nor9, ids, scene["triangleBounds"], n_slabs)


def _build_spheres(ctx, xyzr, ids, bounds, n_slabs):
    xyzr = np.ascontiguousarray(xyzr, dtype=np.float64).reshape(-1, 4)
    ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)
    g = L.Grid()
    ctx.check(L.dll.rt_grid_build_spheres(ctx.h, L.hptr(xyzr), L.hptr(ids), len(xyzr), _d3(bounds.min), _d3(bounds.max), int(n_slabs), C.byref(g)))
    return g


def splitSphereData(ctx, scene, n_slabs) -> "L.Grid":
    """splitSphereData (A10/code.js:1554-1641)."""
    s = scene["spheres"]
    xyzr = np.array([[*x["c"].tolist(), x["r"]] for x in s], dtype=np.float64).reshape(-1, 4)
    ids = np.array([x["matId"] for x in s], dtype=np.uint32)
    return _build_spheres(ctx, xyzr, ids, scene["sphereBounds"], n_slabs)


def splitMolData(ctx, molData, n_slabs) -> "L.Grid":
    """splitMolData (A07/code.js:889-978): visits ``molData.size`` records; records past the
    end of atomData are NaN spheres that land in no cell (Q13)."""
    n = int(molData["size"])
    ad = np.asarray(molData["atomData"], dtype=np.float64).reshape(-1, 4)
    rec = np.full((n, 4), np.nan)
    ids = np.zeros(n, dtype=np.uint32)
    m = min(n, len(ad))
    rec[:m, :3] = ad[:m, 1:4]
    ids[:m] = ad[:m, 0].astype(np.uint32)
    rec[:m, 3] = np.asarray(molData["radiusData"], dtype=np.float64)[ids[:m]]
    return _build_spheres(ctx, rec, ids, molData["bounds"], n_slabs)


def _mol_records(molData):
    n = int(molData["size"])
    ad = np.asarray(molData["atomData"], dtype=np.float64).reshape(-1, 4)
    rec = np.full((n, 4), np.nan)
    ids = np.zeros(n, dtype=np.uint32)
    m = min(n, len(ad))
    rec[:m, :3] = ad[:m, 1:4]
    ids[:m] = ad[:m, 0].astype(np.uint32)
    rec[:m, 3] = np.asarray(molData["radiusData"], dtype=np.float64)[ids[:m]]
    return rec, ids


def slabSplitMolData(ctx, molData, n_slabs) -> "L.Grid":
    """The slab re-ordering inside prepareMolTrace of Assignment 6 (A06/code.js:456-520): atoms binned along x only;
    ``grid.prim`` keeps (x, y, z, RADIUS), ``grid.matid`` the element index (the colour record the reference copies
    per slab reference is ``colorData[4 * matid]``), ``grid.box_size`` the n_slabs + 1 slab limits."""
    rec, ids = _mol_records(molData)
    b = molData["bounds"]
    g = L.Grid()
    ctx.check(L.dll.rt_slab_build_spheres(ctx.h, L.hptr(rec), L.hptr(ids), len(rec), float(b.min[0]), float(b.max[0]), int(n_slabs), C.byref(g)))
    return g


def slabSplitMeshData(ctx, meshData, n_slabs) -> "L.Grid":
    """splitData of Assignment 6 (A06/code.js:936-1043): triangles binned by their x extent."""
    pos9 = np.ascontiguousarray(meshData["positions"], dtype=np.float64).reshape(-1, 9)
    nor9 = np.ascontiguousarray(meshData["normals"], dtype=np.float64).reshape(-1, 9)
    ids = np.ascontiguousarray(meshData["materialIndices"], dtype=np.uint32)
    b = meshData["bounds"]
    g = L.Grid()
    ctx.check(L.dll.rt_slab_build_triangles(ctx.h, L.hptr(pos9), L.hptr(nor9), L.hptr(ids), len(pos9), float(b.min[0]), float(b.max[0]),
                                            int(n_slabs), C.byref(g)))
    return g


def toPosArray(meshData) -> np.ndarray:
    """A04/code.js:845-870: the triangle soup in input order as float4 vertices, w = 1."""
    p = np.asarray(meshData["positions"], dtype=np.float64).reshape(-1, 3)
    out = np.ones((len(p), 4), dtype=np.float32)
    out[:, :3] = p
    return out.reshape(-1)


def toNormalArray(meshData) -> np.ndarray:
    """A04/code.js:819-843: vertex normals in input order, w = 0."""
    p = np.asarray(meshData["normals"], dtype=np.float64).reshape(-1, 3)
    out = np.zeros((len(p), 4), dtype=np.float32)
    out[:, :3] = p
    return out.reshape(-1)


def splitMaterialData(scene) -> np.ndarray:
    """A10/code.js:1774-1782."""
    return np.asarray(scene["materials"], dtype=np.float64).reshape(-1, 4).astype(np.float32)


class Mesh:
    """``Mesh`` (A10/code.js:94-170); normalize -> scale -> translate are each applied at most
    once and in that order, as loadScene does (A10/code.js:866-868).  ``loadFromJSON`` only records the request; the GPU grid
    build runs in ``upload`` once normalize/scale/translate are known, because the reference
    applies them to the cell-ordered positions AFTER the split, in float64, before the fp32
    upload -- the gather kernel does exactly that."""

    def __init__(self):
        self.bounds = Bounds()
        self.ntriangles = 0
        self.nslabs = 1
        self.matId = 0
        self._jmesh = None
        self._split_bounds = None
        self._xform = {"normalize": False, "center": [0.0] * 3, "maxdim": 1.0, "scale": [1.0] * 3, "translate": [0.0] * 3}
        self.grid = None

    def loadFromJSON(self, jmesh, nslabs, matId):
        self.bounds = Bounds(jmesh["bounds"].min, jmesh["bounds"].max)
        self._split_bounds = Bounds(jmesh["bounds"].min, jmesh["bounds"].max)
        self.ntriangles = jmesh["nTriangles"]
        self.nslabs = int(nslabs)
        self.matId = matId
        self._jmesh = jmesh

    def normalize(self):
        mn, mx = self.bounds.min, self.bounds.max
        c = [(mx[a] + mn[a]) / 2.0 for a in range(3)]
        maxdim = 1.0 / max(max(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2])
        self._xform.update(normalize=True, center=c, maxdim=maxdim)
        self.bounds.min = [(mn[a] - c[a]) * maxdim for a in range(3)]
        self.bounds.max = [(mx[a] - c[a]) * maxdim for a in range(3)]

    def scale(self, s):
        f = s.tolist()
        self._xform["scale"] = f
        self.bounds.min = [self.bounds.min[a] * f[a] for a in range(3)]
        self.bounds.max = [self.bounds.max[a] * f[a] for a in range(3)]

    def translate(self, t):
        f = t.tolist()
        self._xform["translate"] = f
        self.bounds.min = [self.bounds.min[a] + f[a] for a in range(3)]
        self.bounds.max = [self.bounds.max[a] + f[a] for a in range(3)]

    def upload(self, ctx):
        if self.grid is None:
            self.grid = splitMeshData(ctx, dict(self._jmesh, bounds=self._split_bounds), self.nslabs, self._xform)
        return self.grid


# ------------------------------------------------------------------------------ XML scenes
def _first(e, name):
    for c in e.iter(name):
        if c is not e:
            return c
    raise KeyError("missing <%s>" % name)


def _num(e, name):
    t = (_first(e, name).text or "").strip()
    return float(t) if t else 0.0


def _str(e, name):
    return _first(e, name).text


def _vec3(e, name):
    v = _first(e, name)
    return Vec3(_num(v, "x"), _num(v, "y"), _num(v, "z"))


def loadScene(sceneName, width, height, mesh_loader=None, assignment=10):
    """loadScene (A10/code.js:723-897; ``assignment`` 8 / 9 select the variants of
    A08/code.js:484-612 and A09/code.js:560-670: point lights that are a bare position, flat
    triangle bounds padded by 1 instead of 0.1, no <mesh>, and no lens parameters in A08).  ``width``/``height`` are the canvas globals.  Files
    start with a UTF-8 BOM and contain commented-out geometry; ``getElementsByTagName`` is a
    descendant search in document order.  ``mesh_loader(file)`` may supply the parseMeshJSON
    result for a ``<mesh>`` (synthetic meshes); by default the path is resolved against the
    assignment directory (the page's base URL)."""
    with open(sceneName, "r", encoding="utf-8-sig") as f:
        doc = ET.fromstring(f.read())
    base = os.path.dirname(os.path.dirname(os.path.abspath(sceneName)))
    xc = _first(doc, "camera")
    cam = Camera()
    cam.lookAt(_vec3(xc, "eye"), _vec3(xc, "lookAt"), _vec3(xc, "vup"), _num(xc, "fov"), width, height)
    focal = lens = None
    if assignment >= 9:
        focal, lens = _num(xc, "focal_length"), _num(xc, "lens_diameter")

    lights = []
    for xl in doc.iter("light"):
        if assignment < 10:
            lights.append(_vec3(xl, "position"))
            continue
        lt = Light()   # fields assigned directly: the light normal is NOT normalised (A10/code.js:751-757)
        lt.position, lt.normal, lt.irradiance = _vec3(xl, "position"), _vec3(xl, "normal"), _vec3(xl, "irradiance")
        lt.radius = _num(xl, "radius")
        lt.calculateArea()
        lt.calculateTBN()
        lights.append(lt)

    materials, lookup = [], {}
    for i, xm in enumerate(doc.iter("material")):
        col = _first(xm, "color")
        materials.append([_num(col, "r"), _num(col, "g"), _num(col, "b"), _num(col, "a")])
        lookup[_str(xm, "id")] = i

    spheres, sphereBounds = [], Bounds()
    for xs in doc.iter("sphere"):
        c, r = _vec3(xs, "center"), _num(xs, "radius")
        spheres.append({"c": c, "r": r, "matId": lookup[_str(xs, "matId")]})
        sphereBounds.merge(Bounds([c.x - r, c.y - r, c.z - r], [c.x + r, c.y + r, c.z + r]))

    triangles, triangleBounds = [], Bounds()
    for xt in doc.iter("triangle"):
        t = {k: _vec3(xt, k) for k in ("p0", "p1", "p2", "n0", "n1", "n2")}
        t["matId"] = lookup[_str(xt, "matId")]
        triangles.append(t)
        ps = [t["p0"].tolist(), t["p1"].tolist(), t["p2"].tolist()]
        triangleBounds.merge(Bounds([min(min(ps[0][a], ps[1][a]), ps[2][a]) for a in range(3)],
                                    [max(max(ps[0][a], ps[1][a]), ps[2][a]) for a in range(3)]))
    for a in range(3):   # zero-thickness guard, A10/code.js:837-842
        if triangleBounds.min[a] == triangleBounds.max[a]:
            pad = 0.1 if assignment >= 10 else 1
            triangleBounds.min[a] -= pad
            triangleBounds.max[a] += pad

    sceneBounds = Bounds()
    meshes = []
    for xm in (doc.iter("mesh") if assignment >= 10 else ()):
        fname = _str(xm, "file")
        jmesh = mesh_loader(fname) if mesh_loader else parseMeshJSON(os.path.join(base, fname))
        mesh = Mesh()
        mesh.loadFromJSON(jmesh, _num(xm, "nslabs"), lookup[_str(xm, "matId")])
        if _str(xm, "normalize") == "yes":
            mesh.normalize()
        mesh.scale(_vec3(xm, "scale"))
        mesh.translate(_vec3(xm, "translate"))
        meshes.append(mesh)
        sceneBounds.merge(mesh.bounds)
    sceneBounds.merge(sphereBounds)
    sceneBounds.merge(triangleBounds)
    return {"camera": cam, "focal_length": focal, "lens_diameter": lens, "lights": lights, "materials": materials,
            "bounds": sceneBounds, "spheres": spheres, "sphereBounds": sphereBounds, "triangles": triangles,
            "triangleBounds": triangleBounds, "meshes": meshes}


def write_png(path, rgba):
    """sendImagetoHTML's canvas.putImageData (A10/code.js:1530-1537) for a headless host: the RGBA image a pass
    returns as an 8-bit PNG (stdlib only)."""
    import struct
    import zlib
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = a.shape[0], a.shape[1]
    raw = b"".join(b"\x00" + a[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6))
                + chunk(b"IEND", b""))


# ------------------------------------------------------------------------------ frame driver
class Renderer:
    """preRender / executeRender / postRender of A10/code.js:1784-1859 on one GPU.

    ``rays_per_pixel``, ``n_slabs`` (global, 1 in A10), ``depth`` (5 in the reference) are the
    JS globals / literals.  ``slots`` = (slot_begin, slot_count) restricts this renderer to a
    sub-range of every pixel's slots (multi-GPU split by samples per pixel)."""

    def __init__(self, scene, width, height, rays_per_pixel=1, n_slabs=1, depth=5, device=0, slots=None, mode=0, tile_slots=0, ctx=None):
        self.scene, self.width, self.height = scene, int(width), int(height)
        self.rays_per_pixel, self.n_slabs, self.depth = int(rays_per_pixel), int(n_slabs), int(depth)
        self.slots = (0, self.rays_per_pixel) if slots is None else (int(slots[0]), int(slots[1]))
        if self.slots[1] <= 0:   # multi.slot_range hands (begin, 0) to ranks beyond rays_per_pixel: they render nothing
            raise ValueError("Renderer: empty slot range %r -- this rank has no slots; skip the render and contribute a zero image" % (self.slots,))
        self.mode, self.tile_slots = mode, tile_slots
        self.ctx = ctx or L.Context(device)
        self._own_ctx = ctx is None
        self.h_scene = self.h_render = None
        self._grids = []
        self.passes = 1

    # -- preRender: A10/code.js:1784-1804 --
    def preRender(self, seeds=None):
        ctx, sc = self.ctx, self.scene
        hs = C.c_void_p()
        ctx.check(L.dll.rt_scene_create(ctx.h, C.byref(hs)))
        self.h_scene = hs
        aabb = bounds2AABB(sc["bounds"])   # keep every host array alive across its C call
        ctx.check(L.dll.rt_scene_set_bounds(hs, L.hptr(aabb)))
        mats = splitMaterialData(sc)
        ctx.check(L.dll.rt_scene_set_materials(hs, L.hptr(mats), len(mats)))
        if len(sc["spheres"]) > 0:
            g = splitSphereData(ctx, sc, self.n_slabs)
            self._grids.append(g)
            aabb = bounds2AABB(sc["sphereBounds"])
            ctx.check(L.dll.rt_scene_add_set(hs, C.byref(g), L.hptr(aabb), 0, 0))
        if len(sc["triangles"]) > 0:
            g = splitTriangleData(ctx, sc, self.n_slabs)
            self._grids.append(g)
            aabb = bounds2AABB(sc["triangleBounds"])
            ctx.check(L.dll.rt_scene_add_set(hs, C.byref(g), L.hptr(aabb), 0, 0))
        for mesh in sc["meshes"]:
            g = mesh.upload(ctx)
            aabb = bounds2AABB(mesh.bounds)
            ctx.check(L.dll.rt_scene_add_set(hs, C.byref(g), L.hptr(aabb), 1, int(mesh.matId)))
        for lt in sc["lights"]:
            i_sh, i_sc, i_li = lt.toShadowInfo(), lt.toSceneRenderInfo(), lt.toLightRenderInfo()
            ctx.check(L.dll.rt_scene_add_light(hs, L.hptr(i_sh), L.hptr(i_sc), L.hptr(i_li)))
        opts = L.RenderOpts(self.width, self.height, self.rays_per_pixel, self.depth, float(np.float32(sc["focal_length"])),
                            float(np.float32(sc["lens_diameter"] / 2.0)), int(self.slots[0]), int(self.slots[1]), int(self.mode), int(self.tile_slots))
        hr = C.c_void_p()
        ctx.check(L.dll.rt_render_create(ctx.h, hs, C.byref(opts), C.byref(hr)))
        self.h_render = hr
        if seeds is not None:
            self.setSeeds(seeds)
        self.passes = 1

    def setSeeds(self, seeds):
        """prepareInitSeeds (A10/code.js:1140-1154) with a caller-supplied Int32Array."""
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        self.ctx.check(L.dll.rt_render_set_seeds(self.h_render, seeds.ctypes.data, seeds.size, 0))

    def writeLocalSeeds(self, seeds, non_blocking=False):
        """Upload THIS renderer's seeds, laid out [pixel][k_local] (= the global layout when it renders every slot).
        ``non_blocking=True`` is the reference's enqueueWriteBuffer(buf, false, ...): the copy overlaps the head of the next
        executeRender (rt_render_write_local_seeds_async); the array is kept alive until that pass has returned."""
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        if non_blocking:
            self._seed_keepalive = seeds
            self.ctx.check(L.dll.rt_render_write_local_seeds_async(self.h_render, seeds.ctypes.data, seeds.size))
        else:
            self.ctx.check(L.dll.rt_render_write_local_seeds(self.h_render, seeds.ctypes.data, seeds.size))

    # -- executeRender: A10/code.js:1806-1854 (one pass, returns the RGBA image) --
    def executeRender(self, camera=None, readback=True):
        cam = (camera or self.scene["camera"]).toFloat32Array()
        img = np.empty((self.height, self.width, 4), dtype=np.uint8) if readback else None
        self.ctx.check(L.dll.rt_render_execute(self.h_render, L.hptr(cam), L.hptr(img)))
        self._seed_keepalive = None
        self.passes += 1
        return img

    def accum(self) -> np.ndarray:
        out = np.empty((self.height * self.width, 4), dtype=np.float32)
        self.ctx.check(L.dll.rt_render_read_accum(self.h_render, out.ctypes.data))
        return out

    def accum_dptr(self) -> int:
        p = C.c_void_p()
        self.ctx.check(L.dll.rt_render_accum_image(self.h_render, C.byref(p)))
        return p.value

    def seeds(self) -> np.ndarray:
        n = self.width * self.height * self.slots[1]
        out = np.empty(n, dtype=np.int32)
        self.ctx.check(L.dll.rt_render_read_seeds(self.h_render, out.ctypes.data, n))
        return out

    def stats(self):
        a, b, l, ms = L.ULL(), L.ULL(), L.U(), L.F()
        self.ctx.check(L.dll.rt_render_stats(self.h_render, C.byref(a), C.byref(b), C.byref(l), C.byref(ms)))
        return {"closest_rays": a.value, "any_rays": b.value, "launches": l.value, "device_ms": ms.value}

    def profile_pass(self, camera=None):
        """Runs ONE extra pass with the instrumented kernel and returns the work counters of
        rt_render_read_profile_sets: one row of 16 counters per geometry set (not a timing run)."""
        self.ctx.check(L.dll.rt_render_set_profile(self.h_render, 1))
        try:
            self.executeRender(camera, readback=False)
            out = (L.ULL * 128)()
            self.ctx.check(L.dll.rt_render_read_profile_sets(self.h_render, C.byref(out)))
        finally:
            self.ctx.check(L.dll.rt_render_set_profile(self.h_render, 0))
        return np.array(list(out), dtype=np.uint64).reshape(8, 16)

    TIMING_CLASSES = ("stage", "walk_sphere_closest", "walk_sphere_any", "walk_triangle_closest", "walk_triangle_any", "megakernel",
                      "reference_schedule", "sum_copy")

    def set_timing(self, on=True):
        """Per-kernel-class CUDA-event timing of the following executeRender calls."""
        self.ctx.check(L.dll.rt_render_set_timing(self.h_render, 1 if on else 0))

    def timing(self):
        ms, n = (L.F * 8)(), (L.U * 8)()
        self.ctx.check(L.dll.rt_render_read_timing(self.h_render, C.byref(ms), C.byref(n)))
        return {k: {"ms": float(ms[i]), "launches": int(n[i])} for i, k in enumerate(self.TIMING_CLASSES)}

    @staticmethod
    def traversal_bytes(p, prim_bytes_sphere=16, prim_bytes_tri=48):
        """ALGORITHMIC traversal bytes of one row of work counters, split into the
        closest-hit and the any-hit queries: 48 B ray load per walk + 8 B per visited cell + 16/48 B
        per sphere/triangle test + per hit the normals (48, triangles), matid (4, not meshes), Poi
        store (64) and maxt store (4); any-hit adds the 8 B mint/maxt store per walk."""
        p = [int(v) for v in p]
        closest = (48 * p[0] + 8 * p[2] + prim_bytes_sphere * p[3] + prim_bytes_tri * p[4] + p[5] * (4 + 64 + 4) + p[6] * (48 + 4 + 64 + 4)
                   + p[7] * (48 + 64 + 4))
        anyhit = 48 * p[8] + 8 * p[10] + prim_bytes_sphere * p[11] + prim_bytes_tri * p[12] + 8 * p[9]
        return closest, anyhit

    def algorithmic_bytes(self, prof=None):
        """ALGORITHMIC bytes of one pass in the REFERENCE's data layout (restated
        in DESIGN.md): what the kernel-by-kernel schedule must move for the work this pass did.
        Streaming kernels, per slot: initTrace 68, lightRender 48/light, bouncePaths 120,
        initShadowTrace 120 and sceneRender 176 per light per segment; copyToPixel 16/slot + 4/px.
        Also returned per geometry set (the queue walkers' share)."""
        p = self.profile_pass() if prof is None else prof
        nl, depth = len(self.scene["lights"]), self.depth
        tot = p.sum(axis=0)
        slots = int(tot[14])
        c, a = self.traversal_bytes(tot)
        stream = slots * (68 + 48 * nl + (1 + depth) * nl * (120 + 176) + depth * 120 + 16) + 4 * self.width * self.height
        per_set = []
        for row in p:
            cs, as_ = self.traversal_bytes(row)
            per_set.append({"closest_bytes": cs, "any_bytes": as_, "closest_walks": int(row[1]), "any_walks": int(row[9]),
                            "closest_queries": int(row[0]), "any_queries": int(row[8]),
                            "cells": int(row[2] + row[10]), "tests": int(row[3] + row[4] + row[11] + row[12])})
        return {"bytes_per_pass": int(c + a + stream), "traversal_bytes": int(c + a), "streaming_bytes": int(stream),
                "profile": [int(v) for v in tot], "per_set": per_set, "profile_sets": [[int(v) for v in row] for row in p]}

    # -- checkpoint / resume of the progressive state (acu, seeds, passes); the reference keeps it only on the device --
    def export_state(self):
        n = self.width * self.height * self.slots[1]
        acu, seeds, passes = np.empty((n, 4), np.float32), np.empty(n, np.int32), L.U()
        self.ctx.check(L.dll.rt_render_export_state(self.h_render, acu.ctypes.data, seeds.ctypes.data, C.byref(passes)))
        return {"acu": acu, "seeds": seeds, "passes": int(passes.value)}

    def import_state(self, state):
        acu = np.ascontiguousarray(state["acu"], dtype=np.float32)
        seeds = np.ascontiguousarray(state["seeds"], dtype=np.int32)
        n = self.width * self.height * self.slots[1]
        if acu.size != 4 * n or seeds.size != n:
            raise ValueError("import_state: state belongs to a render of another size")
        self.ctx.check(L.dll.rt_render_import_state(self.h_render, acu.ctypes.data, seeds.ctypes.data, int(state["passes"])))
        self.passes = int(state["passes"])

    # -- postRender: A10/code.js:1856-1859 --
    def postRender(self):
        if self.h_render:
            L.dll.rt_render_destroy(self.h_render)
            self.h_render = None
        if self.h_scene:
            L.dll.rt_scene_destroy(self.h_scene)
            self.h_scene = None
        for g in self._grids:
            L.dll.rt_grid_release(self.ctx.h, C.byref(g))
        self._grids = []
        for mesh in self.scene["meshes"]:
            if mesh.grid is not None:
                L.dll.rt_grid_release(self.ctx.h, C.byref(mesh.grid))
                mesh.grid = None
        if self._own_ctx:
            self.ctx.close()

    def render(self, seeds):
        """render() (A10/code.js:1883-1894): preRender, one executeRender, postRender."""
        self.preRender(seeds)
        try:
            return self.executeRender()
        finally:
            self.postRender()
