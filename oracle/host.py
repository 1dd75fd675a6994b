"""TEST INFRASTRUCTURE ONLY -- the host half of the parity oracle.

A literal float64 restatement of the reference's JavaScript buffer setup (``code.js`` and
its loaders), statement by statement and in the same operation order, so that the kernel
oracle (``oracle/_ref/libref.so`` = the reference's own ``code.cl`` text, or
``oracle/librt_oracle.so`` = our C restatement) is fed exactly what the browser host
would have uploaded.  JS ``Number`` is an IEEE double and so is a Python ``float``;
``Float32Array`` stores are ``numpy.float32`` casts.

PARITY UNPINNED for this half: the reference's JavaScript cannot be executed in this
image (no browser, Node or JS engine) and the reference ships no test or
golden vector, so this file is pinned only by review against the cited lines.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline/reference
legs may import this module.  All citations are relative to ``/root/reference``
(A10 = Assign10-Path_Tracing, A07 = Assign07-3D_uniform_grid_acceleration, ...).
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

MAX_VALUE = 1.7976931348623157e308  # Number.MAX_VALUE
NAN = float("nan")


# --------------------------------------------------------------------------- JS helpers
def _div(a, b):
    """JS '/' on Numbers (no exception on /0)."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return NAN
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


def _floor(x):
    """Math.floor: NaN and +-inf pass through."""
    if x != x or x in (math.inf, -math.inf):
        return x
    return float(math.floor(x))


def _irange(lo, hi):
    """for (i = lo; i <= hi; i++) with JS Numbers (NaN makes the loop empty)."""
    if lo != lo or hi != hi or lo == math.inf or hi == -math.inf:
        return range(0)
    return range(int(lo), int(hi) + 1)


# --------------------------------------------------------------------------- Bounds
class Bounds:
    """A10/lib/utilities.js:389-422."""

    def __init__(self, mn=None, mx=None):
        self.min = [MAX_VALUE] * 3
        self.max = [-MAX_VALUE] * 3
        if mn is not None:
            self.min = [mn[0], mn[1], mn[2]]
        if mx is not None:
            self.max = [mx[0], mx[1], mx[2]]

    def center(self):
        return [(self.min[i] + self.max[i]) / 2 for i in range(3)]

    def diagonal(self):
        return math.sqrt(
            (self.max[0] - self.min[0]) * (self.max[0] - self.min[0])
            + (self.max[1] - self.min[1]) * (self.max[1] - self.min[1])
            + (self.max[2] - self.min[2]) * (self.max[2] - self.min[2])
        )

    def merge(self, b):
        for i in range(3):
            self.min[i] = min(self.min[i], b.min[i])
        for i in range(3):
            self.max[i] = max(self.max[i], b.max[i])


def bounds2AABB(bounds) -> np.ndarray:
    """A10/code.js:610-621 -- 8 floats, w = 1."""
    a = np.zeros(8, dtype=np.float32)
    with np.errstate(over="ignore"):
        a[0:3] = np.array(bounds.min, dtype=np.float64).astype(np.float32)
        a[3] = 1
        a[4:7] = np.array(bounds.max, dtype=np.float64).astype(np.float32)
        a[7] = 1
    return a


# --------------------------------------------------------------------------- Vec3 & co
class Vec3:
    """A10/code.js:13-53."""

    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z

    def subtract(self, b):
        return Vec3(self.x - b.x, self.y - b.y, self.z - b.z)

    def cross(self, b):
        return Vec3(self.y * b.z - self.z * b.y, self.z * b.x - self.x * b.z, self.x * b.y - self.y * b.x)

    def normalize(self):
        ln = math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)
        self.x = _div(self.x, ln)
        self.y = _div(self.y, ln)
        self.z = _div(self.z, ln)

    def copy(self):
        return Vec3(self.x, self.y, self.z)


class Camera:
    """A10/code.js:175-277 (identical in A07-A09 apart from lookAt's presence)."""

    def __init__(self):
        self.eye, self.U, self.V, self.W = Vec3(), Vec3(), Vec3(), Vec3()
        self.width = 1.0
        self.height = 1.0
        self.cols = 0.0
        self.rows = 0.0

    def defaultInit(self):  # :271-276
        self.eye = Vec3(0.0, 0.0, 0.0)
        self.U = Vec3(1.0, 0.0, 0.0)
        self.V = Vec3(0.0, 1.0, 0.0)
        self.W = Vec3(0.0, 0.0, 1.0)

    def set(self, bounds, cols, rows):  # :185-201
        self.cols, self.rows = cols, rows
        fov = 60
        aspect = cols / rows
        center = bounds.center()
        diag = bounds.diagonal()
        self.eye.x = center[0]
        self.eye.y = center[1]
        self.eye.z = center[2] + diag
        self.height = 2.0 * math.tan(0.5 * fov * math.pi / 180.0)
        self.width = self.height * aspect

    def lookAt(self, eye, lookat, vup, fov, cols, rows):  # :203-217
        self.cols, self.rows = cols, rows
        aspect = cols / rows
        self.height = 2.0 * math.tan(0.5 * fov * math.pi / 180.0)
        self.width = self.height * aspect
        self.eye = eye
        self.W = eye.subtract(lookat)
        self.W.normalize()
        self.U = vup.cross(self.W)
        self.U.normalize()
        self.V = self.W.cross(self.U)

    def rotate(self, bounds, angle):  # :219-245
        center = bounds.center()
        diag = bounds.diagonal()
        rad = angle * math.pi / 180.0
        self.eye.x = center[0] + math.sin(rad) * diag
        self.eye.y = center[1]
        self.eye.z = center[2] + math.cos(rad) * diag
        self.W.x = self.eye.x - center[0]
        self.W.y = self.eye.y - center[1]
        self.W.z = self.eye.z - center[2]
        self.W.normalize()
        self.U = self.V.cross(self.W)

    def toFloat32Array(self):  # :250-258
        return np.array(
            [self.eye.x, self.eye.y, self.eye.z, self.U.x, self.U.y, self.U.z, self.V.x, self.V.y, self.V.z,
             self.W.x, self.W.y, self.W.z, self.width, self.height, self.cols, self.rows], dtype=np.float64
        ).astype(np.float32)


def camera_a01(cols, rows) -> np.ndarray:
    """A01/code.js:43-58,180-185 -- constants; NOTE rows before cols in sE/sF."""
    return np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 2.66, 2.0, rows, cols], dtype=np.float32)


class Light:
    """A10/code.js:279-353 (disk area light)."""

    def __init__(self):
        self.position, self.normal, self.T, self.B, self.irradiance = Vec3(), Vec3(), Vec3(), Vec3(), Vec3()
        self.radius = 0.0
        self.area = 0.0

    def calculateArea(self):  # :298-300
        self.area = math.pi * self.radius * self.radius

    def calculateTBN(self):  # :302-321
        V = Vec3(abs(self.normal.x), abs(self.normal.y), abs(self.normal.z))
        minmag = min(V.x, V.y, V.z)
        if minmag == V.x:
            V = self.normal.copy()
            V.x = 1.0
        elif minmag == V.y:
            V = self.normal.copy()
            V.y = 1.0
        else:
            V = self.normal.copy()
            V.z = 1.0
        V.normalize()
        self.T = V.cross(self.normal)
        self.T.normalize()
        self.B = self.normal.cross(self.T)
        self.B.normalize()

    def _pack(self, a, b, c, s):
        return np.array([a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, s, 0, 0, 0, 0, 0, 0], dtype=np.float64).astype(np.float32)

    def toShadowInfo(self):  # :323-331
        return self._pack(self.position, self.T, self.B, self.radius)

    def toSceneRenderInfo(self):  # :333-342
        return self._pack(self.position, self.normal, self.irradiance, self.area)

    def toLightRenderInfo(self):  # :344-352
        return self._pack(self.position, self.normal, self.irradiance, self.radius)


# --------------------------------------------------------------------------- loaders
def _f32(x):
    return float(np.float32(x))


def parseMeshJSON(path):
    """A10/tri/meshDataVersion1.js:12-78 with gl-matrix 2.2.1 Float32Array semantics
    (A10/lib/gl-matrix.js:79-80, 1063-1085, 2723-2760): matrices and transformed vectors are
    rounded to fp32 on store; the arithmetic itself is double."""
    if isinstance(path, dict):
        model = path
        if sum(len(m["vertexPositions"]) for m in model["meshes"]) // 3 >= FAST_MIN_PRIMS:
            return _parseMeshModel_np(model)
    else:
        with open(path, "r", encoding="utf-8-sig") as f:
            model = json.load(f)
    positions, normals, matidx, materials = [], [], [], []
    b = Bounds()
    nodes = model.get("nodes")
    nNodes = len(nodes) if nodes else 1
    nTriangles = 0
    for k in range(nNodes):
        m = [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0]
        if nodes:
            m = [_f32(v) for v in nodes[k]["modelMatrix"]]
        nm = _normalFromMat4(m)
        idxs = nodes[k]["meshIndices"] if nodes else range(len(model["meshes"]))
        for index in idxs:
            mesh = model["meshes"][index]
            vp, vn = mesh["vertexPositions"], mesh["vertexNormals"]
            for i in range(0, len(vp), 3):
                v = _transformMat4(vp[i], vp[i + 1], vp[i + 2], m)
                for a in range(3):
                    if v[a] < b.min[a]:
                        b.min[a] = v[a]
                    if v[a] > b.max[a]:
                        b.max[a] = v[a]
            ind = mesh.get("indices")
            if ind is not None and len(ind) == 0:
                ind = None
            nV = len(ind) if ind is not None else len(vp) // 3
            nT = nV // 3
            nTriangles += nT
            for i in range(nT):
                for j in range(3):
                    vi = i * 3 + j
                    if ind is not None:
                        vi = ind[vi]
                    positions.extend(_transformMat4(vp[vi * 3], vp[vi * 3 + 1], vp[vi * 3 + 2], m))
                    normals.extend(_transformMat3(vn[vi * 3], vn[vi * 3 + 1], vn[vi * 3 + 2], nm))
                matidx.append(mesh["materialIndex"])
    for mat in model["materials"]:
        materials.extend(mat["diffuseReflectance"][:4])
    return {"nTriangles": nTriangles, "nMaterials": len(model["materials"]), "materialIndices": matidx,
            "materials": materials, "bounds": b, "positions": positions, "normals": normals}


def _parseMeshModel_np(model):
    """Vectorised twin of parseMeshJSON for big in-memory models (same float64 expressions, same
    fp32 rounding points); tests/test_oracle_host.py checks it against the literal loop."""
    def r32(a):
        return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)
    pos, nor, mat = [], [], []
    b = Bounds()
    nodes = model.get("nodes")
    for k in range(len(nodes) if nodes else 1):
        m = [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0]
        if nodes:
            m = [_f32(v) for v in nodes[k]["modelMatrix"]]
        nm = _normalFromMat4(m)
        for index in (nodes[k]["meshIndices"] if nodes else range(len(model["meshes"]))):
            mesh = model["meshes"][index]
            vp = np.asarray(mesh["vertexPositions"], dtype=np.float64).reshape(-1, 3)
            vn = np.asarray(mesh["vertexNormals"], dtype=np.float64).reshape(-1, 3)
            x, y, z = vp[:, 0], vp[:, 1], vp[:, 2]
            tp = r32(np.stack([m[0] * x + m[4] * y + m[8] * z + m[12], m[1] * x + m[5] * y + m[9] * z + m[13],
                               m[2] * x + m[6] * y + m[10] * z + m[14]], axis=1))
            x, y, z = vn[:, 0], vn[:, 1], vn[:, 2]
            tn = r32(np.stack([x * nm[0] + y * nm[3] + z * nm[6], x * nm[1] + y * nm[4] + z * nm[7],
                               x * nm[2] + y * nm[5] + z * nm[8]], axis=1))
            if len(tp):
                for a in range(3):
                    b.min[a] = min(b.min[a], float(tp[:, a].min()))
                    b.max[a] = max(b.max[a], float(tp[:, a].max()))
            ind = mesh.get("indices")
            idx = np.asarray(ind, dtype=np.int64) if (ind is not None and len(ind)) else np.arange(len(vp), dtype=np.int64)
            nT = len(idx) // 3
            idx = idx[:nT * 3]
            pos.append(tp[idx].reshape(-1))
            nor.append(tn[idx].reshape(-1))
            mat.append(np.full(nT, mesh["materialIndex"], dtype=np.int64))
    positions = np.concatenate(pos) if pos else np.zeros(0)
    normals = np.concatenate(nor) if nor else np.zeros(0)
    matidx = np.concatenate(mat) if mat else np.zeros(0, np.int64)
    materials = [c for mt in model["materials"] for c in mt["diffuseReflectance"][:4]]
    return {"nTriangles": len(positions) // 9, "nMaterials": len(model["materials"]), "materialIndices": matidx,
            "materials": materials, "bounds": b, "positions": positions, "normals": normals}


def _transformMat4(x, y, z, m):
    return [_f32(m[0] * x + m[4] * y + m[8] * z + m[12]),
            _f32(m[1] * x + m[5] * y + m[9] * z + m[13]),
            _f32(m[2] * x + m[6] * y + m[10] * z + m[14])]


def _transformMat3(x, y, z, m):
    return [_f32(x * m[0] + y * m[3] + z * m[6]),
            _f32(x * m[1] + y * m[4] + z * m[7]),
            _f32(x * m[2] + y * m[5] + z * m[8])]


def _normalFromMat4(a):
    a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33 = a
    b00 = a00 * a11 - a01 * a10
    b01 = a00 * a12 - a02 * a10
    b02 = a00 * a13 - a03 * a10
    b03 = a01 * a12 - a02 * a11
    b04 = a01 * a13 - a03 * a11
    b05 = a02 * a13 - a03 * a12
    b06 = a20 * a31 - a21 * a30
    b07 = a20 * a32 - a22 * a30
    b08 = a20 * a33 - a23 * a30
    b09 = a21 * a32 - a22 * a31
    b10 = a21 * a33 - a23 * a31
    b11 = a22 * a33 - a23 * a32
    det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06
    if not det:
        raise ValueError("singular modelMatrix (the reference would throw on a null normal matrix)")
    det = 1.0 / det
    out = [
        (a11 * b11 - a12 * b10 + a13 * b09) * det,
        (a12 * b08 - a10 * b11 - a13 * b07) * det,
        (a10 * b10 - a11 * b08 + a13 * b06) * det,
        (a02 * b10 - a01 * b11 - a03 * b09) * det,
        (a00 * b11 - a02 * b08 + a03 * b07) * det,
        (a01 * b08 - a00 * b10 - a03 * b06) * det,
        (a31 * b05 - a32 * b04 + a33 * b03) * det,
        (a32 * b02 - a30 * b05 - a33 * b01) * det,
        (a30 * b04 - a31 * b02 + a33 * b00) * det,
    ]
    return [_f32(v) for v in out]


_ELEMENT_COLORS = {"H": 0xCCCCCC, "C": 0xAAAAAA, "O": 0xCC0000, "N": 0x0000CC, "S": 0xCCCC00, "P": 0x6622CC,
                   "F": 0x00CC00, "CL": 0x00CC00, "BR": 0x882200, "I": 0x6600AA, "FE": 0xCC6600, "CA": 0x8888AA}
_VDW = {"H": 1.2, "Li": 1.82, "Na": 2.27, "K": 2.75, "C": 1.7, "N": 1.55, "O": 1.52, "F": 1.47, "P": 1.80,
        "S": 1.80, "CL": 1.75, "BR": 1.85, "SE": 1.90, "ZN": 1.39, "CU": 1.4, "NI": 1.63}


def _substr(s, a, n):
    return s[a:a + n]


def _parse_float(s):
    s = s.strip()
    try:
        return float(s)
    except ValueError:
        # parseFloat accepts a numeric prefix
        import re
        m = re.match(r"[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", s)
        return float(m.group(0)) if m else NAN


def parsePDB(text):
    """A10/mol/pdbParserV1.js:2-85.  ``atoms[serial-1]`` is a sparse JS array: ``size`` is
    its *length* (max serial), while atomData holds one record per existing atom in index
    order."""
    atoms = {}
    length = 0
    for line in text.split("\n"):
        line = line.lstrip()
        rec = _substr(line, 0, 6)
        if rec == "ATOM  " or rec == "HETATM":
            alt = _substr(line, 16, 1)
            if alt != " " and alt != "A":
                continue
            serial = int(_substr(line, 6, 5))
            x = _parse_float(_substr(line, 30, 8))
            y = _parse_float(_substr(line, 38, 8))
            z = _parse_float(_substr(line, 46, 8))
            elem = _substr(line, 76, 2).replace(" ", "")
            if elem == "":
                elem = _substr(line, 12, 4).replace(" ", "")
            atoms[serial - 1] = (elem, x, y, z)
            length = max(length, serial)
    colorData, radiusData, atomData, used = [], [], [], {}
    minP = [MAX_VALUE] * 3
    maxP = [-MAX_VALUE] * 3
    for i in sorted(atoms):
        elem, x, y, z = atoms[i]
        if elem not in used:
            hexc = _ELEMENT_COLORS[elem]
            colorData.extend([((hexc >> 16) & 255) / 255, ((hexc >> 8) & 255) / 255, (hexc & 255) / 255, 1])
            radiusData.append(_VDW[elem])
            R = _VDW[elem]
            used[elem] = len(used)
            atomData.append(used[elem])
        else:
            atomData.append(used[elem])
            R = radiusData[used[elem]]
        atomData.extend([x, y, z])
        p = (x, y, z)
        for a in range(3):
            if p[a] - R < minP[a]:
                minP[a] = p[a] - R
            if p[a] + R > maxP[a]:
                maxP[a] = p[a] + R
    return {"size": length, "atomData": atomData, "colorData": colorData, "radiusData": radiusData,
            "bounds": Bounds(minP, maxP)}


# --------------------------------------------------------------------------- grid build
def _cell_lists(boxes, bmin, bmax, n):
    """Common tail of split*Data (A10/code.js:940-1002): for each primitive (in input
    order) the inclusive cell box [lo,hi]; append it to every cell, z outer / y / x
    inner; emit exclusive prefix sums over cells in z,y,x order and the cell-ordered list
    of primitive indices (input order preserved inside a cell).
    ``boxes`` yields (min3, max3) per primitive in float64."""
    bw = [_div(bmax[a] - bmin[a], n) for a in range(3)]
    cells = [[] for _ in range(n * n * n)]
    for i, (mn, mx) in enumerate(boxes):
        lo = [_floor(_div(mn[a] - bmin[a], bw[a])) for a in range(3)]
        hi = [_floor(_div(mx[a] - bmin[a], bw[a])) for a in range(3)]
        for a in range(3):
            if lo[a] < 0:
                lo[a] = 0.0
            if hi[a] >= n:
                hi[a] = float(n - 1)
        for iz in _irange(lo[2], hi[2]):
            for iy in _irange(lo[1], hi[1]):
                for ix in _irange(lo[0], hi[0]):
                    cells[(iz * n + iy) * n + ix].append(i)
    box_size = [0]
    total = 0
    order = []
    for c in cells:
        total += len(c)
        box_size.append(total)
        order.extend(c)
    return np.array(box_size, dtype=np.uint32), np.array(order, dtype=np.int64)


FAST_MIN_PRIMS = 20000   # above this the vectorised twins below are used (same results, see tests/test_oracle_host.py)


def _cell_lists_np(mn, mx, bmin, bmax, n):
    """Vectorised twin of _cell_lists for inputs too large for the literal loop: same float64
    operations per primitive ((v - bmin) / bw, floor, one-sided clamps), pairs generated in
    primitive order with x fastest, then a STABLE sort by cell -- i.e. input order inside a cell.
    ``mn``/``mx`` are [N,3] float64.  NaN boxes produce no pairs (JS: the loops are empty)."""
    mn = np.asarray(mn, dtype=np.float64).reshape(-1, 3)
    mx = np.asarray(mx, dtype=np.float64).reshape(-1, 3)
    bmin = np.asarray(bmin, dtype=np.float64)
    bmax = np.asarray(bmax, dtype=np.float64)
    with np.errstate(all="ignore"):
        bw = (bmax - bmin) / float(n)
        lo = np.floor((mn - bmin) / bw)
        hi = np.floor((mx - bmin) / bw)
    bad = np.isnan(lo).any(axis=1) | np.isnan(hi).any(axis=1) | np.isposinf(lo).any(axis=1) | np.isneginf(hi).any(axis=1)
    lo = np.where(lo < 0, 0.0, lo)
    hi = np.where(hi >= n, float(n - 1), hi)
    lo = np.where(np.isneginf(lo), 0.0, lo)
    hi = np.where(np.isposinf(hi), float(n - 1), hi)
    lo[bad] = 1.0
    hi[bad] = 0.0
    loi = lo.astype(np.int64)
    ext = np.maximum(hi.astype(np.int64) - loi + 1, 0)
    cnt = ext[:, 0] * ext[:, 1] * ext[:, 2]
    total = int(cnt.sum())
    off = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    prim = np.repeat(np.arange(len(cnt), dtype=np.int64), cnt)
    j = np.arange(total, dtype=np.int64) - off[prim]
    ex, ey = ext[prim, 0], ext[prim, 1]
    x = loi[prim, 0] + j % np.maximum(ex, 1)
    y = loi[prim, 1] + (j // np.maximum(ex, 1)) % np.maximum(ey, 1)
    z = loi[prim, 2] + j // np.maximum(ex * ey, 1)
    cell = (z * n + y) * n + x
    order = prim[np.argsort(cell, kind="stable")]
    box_size = np.concatenate([[0], np.cumsum(np.bincount(cell, minlength=n * n * n))]).astype(np.uint32)
    return box_size, order


def _tri_boxes_np(pos9):
    p = np.asarray(pos9, dtype=np.float64).reshape(-1, 3, 3)
    return p.min(axis=1), p.max(axis=1)


def _tri_boxes(pos9):
    for i in range(len(pos9) // 9):
        p = pos9[i * 9:i * 9 + 9]
        mn = [min(min(p[a], p[3 + a]), p[6 + a]) for a in range(3)]
        mx = [max(max(p[a], p[3 + a]), p[6 + a]) for a in range(3)]
        yield mn, mx


def _gather_tri(pos9, order):
    """Cell-ordered, w=0-padded float64 [refs*12] (A10/code.js:1003-1034)."""
    p = np.asarray(pos9, dtype=np.float64).reshape(-1, 3, 3)
    out = np.zeros((len(order), 3, 4), dtype=np.float64)
    if len(order):
        out[:, :, :3] = p[order]
    return out.reshape(-1)


def splitMeshData(meshData, nn_slabs):
    """A10/code.js:899-1041 (A07/code.js:980-1122 is the same walk plus a per-reference
    material index).  Returns posData, normalData (float64, before any Mesh transform),
    boxSizeData, indexData."""
    b = meshData["bounds"]
    if len(meshData["positions"]) // 9 >= FAST_MIN_PRIMS:
        box_size, order = _cell_lists_np(*_tri_boxes_np(meshData["positions"]), b.min, b.max, nn_slabs)
    else:
        box_size, order = _cell_lists(_tri_boxes(meshData["positions"]), b.min, b.max, nn_slabs)
    idx = np.asarray(meshData["materialIndices"], dtype=np.uint32)[order] if len(order) else np.zeros(0, np.uint32)
    return _gather_tri(meshData["positions"], order), _gather_tri(meshData["normals"], order), box_size, idx


def splitMolData(molData, n_slabs):
    """A07/code.js:889-978.  Visits ``molData.size`` records -- one more than exist for
    3IZ4.pdb; the missing record is all-NaN and lands in no cell (Q13)."""
    ad, rd = molData["atomData"], molData["radiusData"]
    recs = []
    for i in range(molData["size"]):
        ii = i * 4
        if ii + 3 < len(ad):
            atomId, cx, cy, cz = ad[ii], ad[ii + 1], ad[ii + 2], ad[ii + 3]
            rad = rd[atomId]
        else:
            atomId, cx, cy, cz, rad = 0, NAN, NAN, NAN, NAN
        recs.append((atomId, cx, cy, cz, rad))
    boxes = (([cx - rad, cy - rad, cz - rad], [cx + rad, cy + rad, cz + rad]) for (_, cx, cy, cz, rad) in recs)
    b = molData["bounds"]
    box_size, order = _cell_lists(boxes, b.min, b.max, n_slabs)
    pos = np.zeros((len(order), 4), dtype=np.float64)
    idx = np.zeros(len(order), dtype=np.uint32)
    for k, i in enumerate(order):
        atomId, cx, cy, cz, rad = recs[i]
        pos[k] = (cx, cy, cz, rad * rad)
        idx[k] = atomId
    return pos.reshape(-1), idx, box_size


# --------------------------------------------------------------------------- A04-A06 buffers
def toPosArray(meshData):
    """A04/code.js:845-870 (= A05/code.js:861-886): triangle soup in input order, w = 1."""
    p = np.asarray(meshData["positions"], dtype=np.float64).reshape(-1, 3)
    out = np.ones((len(p), 4), dtype=np.float64)
    out[:, :3] = p
    return out.reshape(-1)


def toNormalArray(meshData):
    """A04/code.js:819-843: vertex normals in input order, w = 0."""
    p = np.asarray(meshData["normals"], dtype=np.float64).reshape(-1, 3)
    out = np.zeros((len(p), 4), dtype=np.float64)
    out[:, :3] = p
    return out.reshape(-1)


def _slab_lists(ranges, x_min, x_max, n_slabs):
    """Common part of A06's two splitters (A06/code.js:456-500, 936-1003): slab k of n along x
    holds every primitive with floor((x_lo - x_min)/w) <= k <= floor((x_hi - x_min)/w), the low
    index clamped at 0 only and the high one at n-1 only; input order inside a slab."""
    w = _div(x_max - x_min, n_slabs)
    slabs = [[] for _ in range(n_slabs)]
    for i, (lo_x, hi_x) in enumerate(ranges):
        lo = _floor(_div(lo_x - x_min, w))
        if lo < 0:
            lo = 0.0
        hi = _floor(_div(hi_x - x_min, w))
        if hi >= n_slabs:
            hi = float(n_slabs - 1)
        for k in _irange(lo, hi):
            slabs[k].append(i)
    limits, order, total = [0], [], 0
    for sl in slabs:
        total += len(sl)
        limits.append(total)
        order.extend(sl)
    return np.array(limits, dtype=np.uint32), np.array(order, dtype=np.int64)


def slabSplitMol(molData, n_slabs):
    """prepareMolTrace of A06 (A06/code.js:456-520): atoms re-ordered into x slabs.  Returns
    atomData (x,y,z,RADIUS -- not squared) and colorData per slab reference (float64, before the
    Float32Array store), slab limits, and the element index per reference.  Records past the end
    of atomData (3IZ4.pdb, Q13) read `undefined`, bin to NaN and land in no slab."""
    ad, rd, cd = molData["atomData"], molData["radiusData"], molData["colorData"]
    recs = []
    for i in range(molData["size"]):
        ii = i * 4
        if ii + 3 < len(ad):
            atomId, cx, cy, cz = ad[ii], ad[ii + 1], ad[ii + 2], ad[ii + 3]
            rad = rd[atomId]
        else:
            atomId, cx, cy, cz, rad = 0, NAN, NAN, NAN, NAN
        recs.append((atomId, cx, cy, cz, rad))
    b = molData["bounds"]
    limits, order = _slab_lists(((cx - rad, cx + rad) for (_, cx, _cy, _cz, rad) in recs), b.min[0], b.max[0], n_slabs)
    atoms = np.zeros((len(order), 4), dtype=np.float64)
    colors = np.zeros((len(order), 4), dtype=np.float64)
    idx = np.zeros(len(order), dtype=np.uint32)
    for k, i in enumerate(order):
        atomId, cx, cy, cz, rad = recs[i]
        atoms[k] = (cx, cy, cz, rad)
        colors[k] = cd[atomId * 4:atomId * 4 + 4]
        idx[k] = atomId
    return atoms.reshape(-1), colors.reshape(-1), limits, idx


def slabSplitMesh(meshData, n_slabs):
    """splitData of A06 (A06/code.js:936-1043).  Returns posData, normalData (w = 0 padded,
    float64), indexData, slabSizeData."""
    pos = np.asarray(meshData["positions"], dtype=np.float64).reshape(-1, 3, 3)
    xs = pos[:, :, 0]
    ranges = ((min(min(x[0], x[1]), x[2]), max(max(x[0], x[1]), x[2])) for x in xs.tolist())
    b = meshData["bounds"]
    limits, order = _slab_lists(ranges, b.min[0], b.max[0], n_slabs)
    idx = np.asarray(meshData["materialIndices"], dtype=np.uint32)[order] if len(order) else np.zeros(0, np.uint32)
    return _gather_tri(meshData["positions"], order), _gather_tri(meshData["normals"], order), idx, limits


def splitSphereData(scene, n_slabs):
    """A10/code.js:1554-1641."""
    sph = scene["spheres"]
    boxes = (([s["c"].x - s["r"], s["c"].y - s["r"], s["c"].z - s["r"]],
              [s["c"].x + s["r"], s["c"].y + s["r"], s["c"].z + s["r"]]) for s in sph)
    b = scene["sphereBounds"]
    box_size, order = _cell_lists(boxes, b.min, b.max, n_slabs)
    data = np.zeros((len(order), 4), dtype=np.float64)
    mat = np.zeros(len(order), dtype=np.uint32)
    for k, i in enumerate(order):
        s = sph[i]
        data[k] = (s["c"].x, s["c"].y, s["c"].z, s["r"] * s["r"])
        mat[k] = s["matId"]
    return data.reshape(-1), mat, box_size


def splitTriangleData(scene, n_slabs):
    """A10/code.js:1643-1772."""
    tris = scene["triangles"]
    pos9, nor9 = [], []
    for t in tris:
        for p in (t["p0"], t["p1"], t["p2"]):
            pos9.extend([p.x, p.y, p.z])
        for nn in (t["n0"], t["n1"], t["n2"]):
            nor9.extend([nn.x, nn.y, nn.z])
    b = scene["triangleBounds"]
    box_size, order = _cell_lists(_tri_boxes(pos9), b.min, b.max, n_slabs)
    mat = np.array([tris[i]["matId"] for i in order], dtype=np.uint32)
    return _gather_tri(pos9, order), _gather_tri(nor9, order), mat, box_size


class Mesh:
    """A10/code.js:94-170.  Transforms act on the cell-ordered posData AFTER the grid
    build, in float64; the bounds are transformed alike; normals are untouched."""

    def __init__(self):
        self.posData = np.zeros(0)
        self.normalData = np.zeros(0)
        self.boxSizeData = np.zeros(1, np.uint32)
        self.bounds = Bounds()
        self.ntriangles = 0
        self.nslabs = 1
        self.matId = 0

    def loadFromJSON(self, jmesh, nslabs, matId):
        self.bounds = jmesh["bounds"]
        self.ntriangles = jmesh["nTriangles"]
        self.nslabs = nslabs
        self.posData, self.normalData, self.boxSizeData, _ = splitMeshData(jmesh, nslabs)
        self.matId = matId

    def normalize(self):
        bmin, bmax = self.bounds.min, self.bounds.max
        c = [(bmax[a] + bmin[a]) / 2.0 for a in range(3)]
        dim = [bmax[a] - bmin[a] for a in range(3)]
        maxdim = 1.0 / max(max(dim[0], dim[1]), dim[2])
        p = self.posData.reshape(-1, 4)
        for a in range(3):
            p[:, a] = (p[:, a] - c[a]) * maxdim
        for a in range(3):
            bmin[a] = (bmin[a] - c[a]) * maxdim
        for a in range(3):
            bmax[a] = (bmax[a] - c[a]) * maxdim

    def scale(self, s):
        p = self.posData.reshape(-1, 4)
        for a, f in enumerate((s.x, s.y, s.z)):
            p[:, a] *= f
            self.bounds.min[a] *= f
            self.bounds.max[a] *= f

    def translate(self, t):
        p = self.posData.reshape(-1, 4)
        for a, f in enumerate((t.x, t.y, t.z)):
            p[:, a] += f
            self.bounds.min[a] += f
            self.bounds.max[a] += f


# --------------------------------------------------------------------------- XML scenes
def _num(s):
    """JS Number(string)."""
    s = s.strip()
    return float(s) if s else 0.0


def _first(elem, name):
    """getElementsByTagName(name)[0] -- first DESCENDANT in document order."""
    for e in elem.iter(name):
        if e is not elem:
            return e
    raise KeyError(name)


def _xml_vec3(elem, name):
    e = _first(elem, name)
    return Vec3(_num(_first(e, "x").text), _num(_first(e, "y").text), _num(_first(e, "z").text))


def _xml_num(elem, name):
    return _num(_first(elem, name).text)


def _xml_str(elem, name):
    return _first(elem, name).text


def _tri_bounds(t):
    mn = [min(min(getattr(t["p0"], a), getattr(t["p1"], a)), getattr(t["p2"], a)) for a in "xyz"]
    mx = [max(max(getattr(t["p0"], a), getattr(t["p1"], a)), getattr(t["p2"], a)) for a in "xyz"]
    return Bounds(mn, mx)


def loadScene(path, width, height, assignment=10, mesh_loader=None):
    """A10/code.js:723-897; ``assignment`` 8/9 selects the A08/A09 variants
    (A08/code.js:484-612: point lights, +-1 flat-bounds padding, no meshes, no lens).
    ``mesh_loader(file) -> parseMeshJSON dict`` lets tests substitute synthetic meshes;
    by default ``<file>`` is resolved relative to the assignment directory."""
    with open(path, "r", encoding="utf-8-sig") as f:
        root = ET.fromstring(f.read())  # comments are dropped by the parser, as in the DOM walk
    base = os.path.dirname(os.path.dirname(os.path.abspath(path)))
    sceneBounds = Bounds()
    cam = Camera()
    xc = _first(root, "camera")
    cam.lookAt(_xml_vec3(xc, "eye"), _xml_vec3(xc, "lookAt"), _xml_vec3(xc, "vup"), _xml_num(xc, "fov"), width, height)
    f_len = ld = None
    if assignment >= 9:
        f_len = _xml_num(xc, "focal_length")
        ld = _xml_num(xc, "lens_diameter")

    lights = []
    for xl in root.iter("light"):
        if assignment >= 10:
            L = Light()
            L.position = _xml_vec3(xl, "position")
            L.normal = _xml_vec3(xl, "normal")  # NOT normalised on load (Q4)
            L.irradiance = _xml_vec3(xl, "irradiance")
            L.radius = _xml_num(xl, "radius")
            L.calculateArea()
            L.calculateTBN()
            lights.append(L)
        else:
            lights.append(_xml_vec3(xl, "position"))

    materials, lookup = [], {}
    for i, xm in enumerate(root.iter("material")):
        xcol = _first(xm, "color")
        materials.append((_xml_num(xcol, "r"), _xml_num(xcol, "g"), _xml_num(xcol, "b"), _xml_num(xcol, "a")))
        lookup[_xml_str(xm, "id")] = i

    spheres, sphereBounds = [], Bounds()
    for xs in root.iter("sphere"):
        s = {"c": _xml_vec3(xs, "center"), "r": _xml_num(xs, "radius"), "matId": lookup[_xml_str(xs, "matId")]}
        sphereBounds.merge(Bounds([s["c"].x - s["r"], s["c"].y - s["r"], s["c"].z - s["r"]],
                                  [s["c"].x + s["r"], s["c"].y + s["r"], s["c"].z + s["r"]]))
        spheres.append(s)

    triangles, triangleBounds = [], Bounds()
    for xt in root.iter("triangle"):
        t = {k: _xml_vec3(xt, k) for k in ("p0", "p1", "p2", "n0", "n1", "n2")}
        t["matId"] = lookup[_xml_str(xt, "matId")]
        triangleBounds.merge(_tri_bounds(t))
        triangles.append(t)
    pad = 0.1 if assignment >= 10 else 1.0  # A10/code.js:837-842 ; A08/code.js:590-594
    for a in range(3):
        if triangleBounds.min[a] == triangleBounds.max[a]:
            triangleBounds.min[a] -= pad
            triangleBounds.max[a] += pad

    meshes = []
    if assignment >= 10:
        for xm in root.iter("mesh"):
            fmesh = _xml_str(xm, "file")
            nslabs = int(_xml_num(xm, "nslabs"))
            normalize = _xml_str(xm, "normalize") == "yes"
            scale = _xml_vec3(xm, "scale")
            translate = _xml_vec3(xm, "translate")
            matId = lookup[_xml_str(xm, "matId")]
            jmesh = mesh_loader(fmesh) if mesh_loader else parseMeshJSON(os.path.join(base, fmesh))
            mesh = Mesh()
            mesh.loadFromJSON(jmesh, nslabs, matId)
            if normalize:
                mesh.normalize()
            mesh.scale(scale)
            mesh.translate(translate)
            meshes.append(mesh)
            sceneBounds.merge(mesh.bounds)
    sceneBounds.merge(sphereBounds)
    sceneBounds.merge(triangleBounds)
    return {"camera": cam, "focal_length": f_len, "lens_diameter": ld, "lights": lights, "materials": materials,
            "bounds": sceneBounds, "spheres": spheres, "sphereBounds": sphereBounds, "triangles": triangles,
            "triangleBounds": triangleBounds, "meshes": meshes}


def splitMaterialData(scene) -> np.ndarray:
    """A10/code.js:1774-1782."""
    return np.array(scene["materials"], dtype=np.float64).reshape(-1, 4).astype(np.float32)


def to_f32(a) -> np.ndarray:
    """``new Float32Array(jsArray)``: round-to-nearest double -> float store."""
    with np.errstate(over="ignore", invalid="ignore"):
        return np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32))
