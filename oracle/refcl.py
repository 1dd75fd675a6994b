"""TEST INFRASTRUCTURE ONLY -- ctypes access to the kernel oracles and the reference's
launch schedules.

Two interchangeable kernel back ends export the same C entry points (one per reference
kernel, arguments in the kernel's own order):

* ``oracle/_ref/libref.so``   prefix ``ref_``  -- the reference's own ``code.cl`` text compiled
  by g++ behind ``clshim.h`` (kind "reference"); ``libref_instr.so`` adds counters.
* ``oracle/librt_oracle.so``  prefix ``port_`` -- our plain-C restatement (kind "port").

The frame schedules below restate ``executeRender``/``render``/``compute*`` of each
assignment's ``code.js`` (cited per function).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import host as H

HERE = os.path.dirname(os.path.abspath(__file__))

RAY = np.dtype([("o", "f4", 4), ("d", "f4", 4), ("mint", "f4"), ("maxt", "f4"), ("pad", "f4", 2)])
POI10 = np.dtype([("p", "f4", 4), ("normal", "f4", 4), ("atte", "f4", 4), ("matId", "i4"), ("pad", "i4", 3)])
POI8 = np.dtype([("p", "f4", 4), ("normal", "f4", 4), ("matId", "i4"), ("pad", "i4", 3)])
assert RAY.itemsize == 48 and POI10.itemsize == 64 and POI8.itemsize == 48

_P = C.c_void_p
_U = C.c_uint
_F = C.c_float
_I = C.c_int

# name -> argtypes (shared by both back ends)
_SIGS = {
    "a01_raytrace": [_P, _P, _U, _U],
    "a02_raytrace": [_P, _P, _U, _P, _P, _U, _U],
    "a03_initTrace": [_P, _P, _P, _U, _U],
    "a03_molTrace": [_P, _P, _P, _U, _P, _P, _U, _U],
    "a04_initTrace": [_P, _P, _P, _U, _U],
    "a04_molTrace": [_P, _P, _P, _U, _P, _P, _U, _U],
    "a04_meshTrace": [_P, _P, _P, _U, _P, _P, _P, _P, _U, _U],
    "a04_raytrace": [_P, _P, _U, _P, _P, _U, _U],
    "a05_initTrace": [_P, _P, _P, _P, _U, _U],
    "a05_molTrace": [_P, _P, _P, _U, _P, _P, _P, _U, _U],
    "a05_meshTrace": [_P, _P, _P, _U, _P, _P, _P, _P, _P, _U, _U],
    "a06_initTrace": [_P, _P, _P, _P, _U, _U],
    "a06_molTrace": [_P, _P, _P, _U, _P, _P, _P, _U, _P, _U, _U],
    "a06_meshTrace": [_P, _P, _P, _U, _P, _P, _P, _P, _P, _U, _P, _U, _U],
    "a07_initTrace": [_P, _P, _P, _P, _U, _U],
    "a07_molTrace": [_P, _P, _P, _U, _P, _P, _P, _P, _U, _P, _U, _U],
    "a07_meshTrace": [_P, _P, _P, _U, _P, _P, _P, _P, _P, _U, _P, _U, _U],
    "a08_initTrace": [_P, _P, _P, _P, _P, _U, _U],
    "a08_initShadowTrace": [_P, _P, _U, _U, _P],
    "a08_sphereTrace": [_U, _U, _P, _P, _P, _P, _P, _P, _U],
    "a08_triangleTrace": [_U, _U, _P, _P, _P, _P, _P, _P, _P, _U],
    "a08_sphereShadowTrace": [_U, _U, _P, _P, _P, _P, _U],
    "a08_triangleShadowTrace": [_U, _U, _P, _P, _P, _P, _U],
    "a08_sceneRender": [_P, _P, _P, _P, _U],
    "a08_copyToPixel": [_P, _P, _F, _U],
    "a09_initTrace": [_P, _P, _P, _P, _P, _F, _F, _U, _U, _U],
    "a09_initShadowTrace": [_P, _P, _U, _P],
    "a09_sphereTrace": [_U, _P, _P, _P, _P, _P, _P, _U],
    "a09_triangleTrace": [_U, _P, _P, _P, _P, _P, _P, _P, _U],
    "a09_sphereShadowTrace": [_U, _P, _P, _P, _P, _U],
    "a09_triangleShadowTrace": [_U, _P, _P, _P, _P, _U],
    "a09_sceneRender": [_P, _P, _P, _P, _U],
    "a09_copyToPixel": [_P, _P, _F, _U, _U],
    "a10_initAcu": [_P, _U],
    "a10_initTrace": [_P, _P, _P, _P, _P, _F, _F, _U, _U, _U, _I],
    "a10_initTrace_rows": [_P, _P, _P, _P, _P, _F, _F, _U, _U, _U, _U, _U],
    "a10_bouncePaths": [_P, _P, _P, _U],
    "a10_lightRender": [_P, _P, _P, _P, _U],
    "a10_initShadowTrace": [_P, _P, _U, _P, _P],
    "a10_sphereTrace": [_U, _P, _P, _P, _P, _P, _P, _U],
    "a10_triangleTrace": [_U, _P, _P, _P, _P, _P, _P, _P, _U],
    "a10_meshTrace": [_U, _P, _P, _P, _P, _P, _U, _P, _U],
    "a10_sphereShadowTrace": [_U, _P, _P, _P, _P, _U],
    "a10_triangleShadowTrace": [_U, _P, _P, _P, _P, _U],
    "a10_sceneRender": [_P, _P, _P, _P, _P, _U],
    "a10_copyToPixel": [_P, _P, _F, _U, _U],
}
_SIZEOF = ["a03_sizeofRay", "a04_sizeofRay", "a05_sizeofRay", "a06_sizeofRay", "a07_sizeofRay", "a08_sizeofRay", "a08_sizeofPoi", "a09_sizeofRay", "a09_sizeofPoi",
           "a10_sizeofRay", "a10_sizeofPoi"]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "oracle buffers must be contiguous"
        return a.ctypes.data
    return a


class KernelLib:
    """One kernel back end.  ``lib.a10_sphereTrace(total_rays, pois, rays, ...)`` with numpy
    arrays for pointer arguments."""

    def __init__(self, path, prefix, kind):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build it with `make -C oracle`)")
        self.path, self.prefix, self.kind = path, prefix, kind
        self._dll = C.CDLL(path)
        for name, sig in _SIGS.items():
            fn = getattr(self._dll, prefix + name, None)
            if fn is None:   # the plain-C restatement covers the Assignment-10 kernels only
                setattr(self, name, self._missing(name))
                continue
            fn.argtypes = sig
            fn.restype = None
            setattr(self, name, self._wrap(fn))
        for name in _SIZEOF:
            fn = getattr(self._dll, prefix + name, None)
            if fn is None:
                setattr(self, name, self._missing(name))
                continue
            fn.argtypes = []
            fn.restype = _U
            setattr(self, name, fn)
        self.num_threads = getattr(self._dll, prefix + "num_threads")
        self.num_threads.restype = _I
        self.set_num_threads = getattr(self._dll, prefix + "set_num_threads", lambda n: None)
        self._set_stats = getattr(self._dll, prefix + "set_stats", None)
        if self._set_stats is not None:
            self._set_stats.argtypes = [_P, _P, _P]
        self._is_instr = getattr(self._dll, prefix + "is_instrumented", None)

    def _missing(self, name):
        def call(*_a):
            raise NotImplementedError("%s is not provided by the %s oracle (%s)" % (name, self.kind, self.path))
        return call

    @staticmethod
    def _wrap(fn):
        def call(*args):
            return fn(*[_ptr(a) for a in args])
        return call

    @property
    def instrumented(self):
        return bool(self._is_instr and self._is_instr())

    def set_stats(self, hit_id=None, cells=None, tests=None):
        """Per-work-item sinks (uint32 / uint64 / uint64 arrays or None) filled by the next
        launches -- only meaningful on an instrumented build."""
        self._set_stats(_ptr(hit_id), _ptr(cells), _ptr(tests))


def load_reference(instrumented=False) -> KernelLib:
    name = "libref_instr.so" if instrumented else "libref.so"
    return KernelLib(os.path.join(HERE, "_ref", name), "ref_", "reference")


def load_port() -> KernelLib:
    return KernelLib(os.path.join(HERE, "librt_oracle.so"), "port_", "port")


def have_reference() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref.so"))


def load_best() -> KernelLib:
    """The reference's own text when it was compiled here, else our C restatement."""
    return load_reference() if have_reference() else load_port()


# ===================================================================================
# Host-side buffer preparation ("what the browser host uploads")
# ===================================================================================
def prepare_a10(scene, n_slabs=1):
    """A10/code.js:1156-1291, 1364-1395 -- Float32Array/Uint32Array conversions of the
    split*Data outputs, AABBs via bounds2AABB, light packings."""
    out = {"aabb": H.bounds2AABB(scene["bounds"]), "materials": H.splitMaterialData(scene), "sets": []}
    if len(scene["spheres"]) > 0:
        data, mat, box = H.splitSphereData(scene, n_slabs)
        out["sets"].append({"kind": "sphere", "data": H.to_f32(data), "matid": mat.astype(np.uint32), "box": box,
                            "aabb": H.bounds2AABB(scene["sphereBounds"]), "n": n_slabs})
    if len(scene["triangles"]) > 0:
        pos, nor, mat, box = H.splitTriangleData(scene, n_slabs)
        out["sets"].append({"kind": "triangle", "pos": H.to_f32(pos), "normal": H.to_f32(nor), "matid": mat.astype(np.uint32),
                            "box": box, "aabb": H.bounds2AABB(scene["triangleBounds"]), "n": n_slabs})
    for m in scene["meshes"]:
        out["sets"].append({"kind": "mesh", "pos": H.to_f32(m.posData), "normal": H.to_f32(m.normalData),
                            "matid": int(m.matId), "box": np.asarray(m.boxSizeData, dtype=np.uint32),
                            "aabb": H.bounds2AABB(m.bounds), "n": int(m.nslabs)})
    out["lights"] = [{"shadow": L.toShadowInfo(), "scene": L.toSceneRenderInfo(), "light": L.toLightRenderInfo()}
                     for L in scene["lights"]]
    return out


def make_seeds(total_rays, seed=2015) -> np.ndarray:
    """Stand-in for ``1 + Math.floor(Math.random()*2147483647)`` (A10/code.js:1140-1146):
    documented generator so oracle and product see the same array."""
    return np.random.Generator(np.random.PCG64(seed)).integers(1, 2 ** 31, size=total_rays, dtype=np.int64).astype(np.int32)


class A10State:
    """Device-side state of an A10 render: rays, pois, shadow rays, acu, seeds."""

    def __init__(self, total_rays, seeds):
        self.total = total_rays
        self.rays = np.zeros(total_rays, dtype=RAY)
        self.pois = np.zeros(total_rays, dtype=POI10)
        self.shadow = np.zeros(total_rays, dtype=RAY)
        self.acu = np.zeros((total_rays, 4), dtype=np.float32)
        self.seeds = np.ascontiguousarray(seeds, dtype=np.int32).copy()
        self.passes = 1
        self.n_closest = 0   # valid closest-hit queries (rays) traced so far
        self.n_any = 0       # valid any-hit queries traced so far


def _count_valid(rays):
    return int(np.count_nonzero(rays["mint"] != rays["maxt"]))


def _closest(lib, st, prep):
    st.n_closest += _count_valid(st.rays)
    for s in prep["sets"]:
        if s["kind"] == "sphere":
            lib.a10_sphereTrace(st.total, st.pois, st.rays, s["data"], s["matid"], s["box"], s["aabb"], s["n"])
        elif s["kind"] == "triangle":
            lib.a10_triangleTrace(st.total, st.pois, st.rays, s["pos"], s["normal"], s["matid"], s["box"], s["aabb"], s["n"])
        else:
            lib.a10_meshTrace(st.total, st.pois, st.rays, s["pos"], s["normal"], s["box"], s["matid"], s["aabb"], s["n"])


def _shade(lib, st, prep, light):
    lib.a10_initShadowTrace(st.shadow, st.pois, st.total, light["shadow"], st.seeds)
    st.n_any += _count_valid(st.shadow)
    for s in prep["sets"]:
        if s["kind"] == "sphere":
            lib.a10_sphereShadowTrace(st.total, st.shadow, s["data"], s["box"], s["aabb"], s["n"])
        else:
            lib.a10_triangleShadowTrace(st.total, st.shadow, s["pos"], s["box"], s["aabb"], s["n"])
    lib.a10_sceneRender(st.acu, st.pois, st.shadow, prep["materials"], light["scene"], st.total)


def a10_execute_render(lib, st, prep, cam16, cols, rows, rpp, focal_length, lens_diameter, depth=5, serial_init=False,
                       row0=0, nrows=None):
    """One pass = A10/code.js:1806-1854 (executeRender).  Returns the uchar4 image.
    ``row0``/``nrows`` restrict the pass to a tile of pixel rows (``st`` then holds only the
    tile's slots) -- exact, because every kernel touches only its own slot."""
    lens_rad = float(np.float32(lens_diameter / 2.0))
    if nrows is None:
        lib.a10_initTrace(st.seeds, st.rays, st.pois, prep["aabb"], cam16, float(np.float32(focal_length)), lens_rad, rpp,
                          cols, rows, 1 if (serial_init or rpp == 1) else 0)
    else:
        assert rpp > 1, "row tiles need the stratified (RNG-free) initTrace"
        lib.a10_initTrace_rows(st.seeds, st.rays, st.pois, prep["aabb"], cam16, float(np.float32(focal_length)), lens_rad, rpp,
                               cols, rows, row0, nrows)
        rows = nrows
    _closest(lib, st, prep)
    for L in prep["lights"]:
        lib.a10_lightRender(st.pois, st.rays, st.acu, L["light"], st.total)
    for L in prep["lights"]:
        _shade(lib, st, prep, L)
    for _ in range(depth):
        lib.a10_bouncePaths(st.pois, st.rays, st.seeds, st.total)
        _closest(lib, st, prep)
        for L in prep["lights"]:
            _shade(lib, st, prep, L)
    pixel = np.zeros((rows * cols, 4), dtype=np.uint8)
    m = float(np.float32(1.0 / (rpp * st.passes)))
    lib.a10_copyToPixel(pixel, st.acu, m, cols * rows, rpp)
    st.passes += 1
    return pixel.reshape(rows, cols, 4)


def a10_render(lib, scene, cols, rows, rpp, passes=1, seeds=None, seed=2015, n_slabs=1, depth=5):
    """preRender + ``passes`` x executeRender (A10/code.js:1784-1804, 1861-1881)."""
    prep = prepare_a10(scene, n_slabs)
    total = rpp * cols * rows
    if seeds is None:
        seeds = make_seeds(total, seed)
    st = A10State(total, seeds)
    lib.a10_initAcu(st.acu, total)
    cam16 = scene["camera"].toFloat32Array()
    pixel = None
    for _ in range(passes):
        pixel = a10_execute_render(lib, st, prep, cam16, cols, rows, rpp, scene["focal_length"], scene["lens_diameter"], depth)
    return st, pixel, prep


# ===================================================================================
# Frame schedules of the earlier assignments (BASELINE.json configs 1-4)
# ===================================================================================
def a01_render(lib, cols, rows):
    """A01/code.js:166-269 (compute): constant camera, one `raytrace` launch."""
    pix = np.zeros((rows * cols, 4), dtype=np.uint8)
    lib.a01_raytrace(pix, H.camera_a01(cols, rows), cols, rows)
    return pix.reshape(rows, cols, 4)


def pack_atoms(molData):
    """A03/code.js:524-537: Float32Array(size*4) atom (x,y,z,radius) and colour records; records
    past the end of atomData read `undefined` and become NaN (Q13)."""
    n = int(molData["size"])
    ad = np.asarray(molData["atomData"], dtype=np.float64).reshape(-1, 4)
    rd = np.asarray(molData["radiusData"], dtype=np.float64)
    cd = np.asarray(molData["colorData"], dtype=np.float64).reshape(-1, 4)
    atoms = np.full((n, 4), np.nan)
    colors = np.full((n, 4), np.nan)
    m = min(n, len(ad))
    ids = ad[:m, 0].astype(np.int64)
    atoms[:m, :3] = ad[:m, 1:4]
    atoms[:m, 3] = rd[ids]
    colors[:m] = cd[ids]
    return H.to_f32(atoms.reshape(-1)), H.to_f32(colors.reshape(-1))


def mol_camera(bounds, cols, rows):
    """cam.defaultInit(); cam.set(bounds, width, height) -- A03/code.js:55-70,113-118,
    A07/code.js:45-124."""
    cam = H.Camera()
    cam.defaultInit()
    cam.set(bounds, cols, rows)
    return cam


def a02_render(lib, molData, cols, rows):
    """A02/code.js:413-573: one fused `raytrace` launch over all atoms."""
    atoms, colors = pack_atoms(molData)
    cam16 = mol_camera(molData["bounds"], cols, rows).toFloat32Array()
    pix = np.zeros((rows * cols, 4), dtype=np.uint8)
    lib.a02_raytrace(pix, cam16, int(molData["size"]), atoms, colors, cols, rows)
    return pix.reshape(rows, cols, 4)


def a03_render(lib, molData, cols, rows):
    """A03/code.js:450-598: `initTrace` then `molTrace`.  Returns (pixels, rays)."""
    atoms, colors = pack_atoms(molData)
    cam16 = mol_camera(molData["bounds"], cols, rows).toFloat32Array()
    pix = np.zeros((rows * cols, 4), dtype=np.uint8)
    rays = np.zeros(rows * cols, dtype=RAY)
    lib.a03_initTrace(pix, cam16, rays, cols, rows)
    lib.a03_molTrace(pix, cam16, rays, int(molData["size"]), atoms, colors, cols, rows)
    return pix.reshape(rows, cols, 4), rays


def _merged_bounds(molData, meshData):
    if molData is not None and meshData is not None:
        bounds = H.Bounds()
        bounds.merge(molData["bounds"])
        bounds.merge(meshData["bounds"])
        return bounds
    return (molData or meshData)["bounds"]


def prepare_a045_mesh(meshData):
    """prepareMeshTrace of A04/A05 (A04/code.js:448-492): the triangle soup in input order."""
    return {"size": int(meshData["nTriangles"]), "pos": H.to_f32(H.toPosArray(meshData)), "normal": H.to_f32(H.toNormalArray(meshData)),
            "index": np.asarray(meshData["materialIndices"], dtype=np.uint32), "colors": H.to_f32(meshData["materials"]),
            "aabb": H.bounds2AABB(meshData["bounds"])}


def a04_render(lib, cols, rows, molData=None, meshData=None):
    """compute / computeTri / computeBoth of A04 (A04/code.js:520-606): `initTrace`, then brute-force
    `molTrace` and/or `meshTrace` over one ray buffer.  Returns (pixels, rays)."""
    cam16 = mol_camera(_merged_bounds(molData, meshData), cols, rows).toFloat32Array()
    pix = np.zeros((rows * cols, 4), dtype=np.uint8)
    rays = np.zeros(rows * cols, dtype=RAY)
    lib.a04_initTrace(pix, cam16, rays, cols, rows)
    if molData is not None:
        atoms, colors = pack_atoms(molData)
        lib.a04_molTrace(pix, cam16, rays, int(molData["size"]), atoms, colors, cols, rows)
    if meshData is not None:
        t = prepare_a045_mesh(meshData)
        lib.a04_meshTrace(pix, cam16, rays, t["size"], t["pos"], t["normal"], t["index"], t["colors"], cols, rows)
    return pix.reshape(rows, cols, 4), rays


def a04_raytrace(lib, molData, cols, rows):
    """The fused `raytrace` kernel A04 still carries (A04/code.cl:317-364; not launched by its code.js)."""
    atoms, colors = pack_atoms(molData)
    cam16 = mol_camera(molData["bounds"], cols, rows).toFloat32Array()
    pix = np.zeros((rows * cols, 4), dtype=np.uint8)
    lib.a04_raytrace(pix, cam16, int(molData["size"]), atoms, colors, cols, rows)
    return pix.reshape(rows, cols, 4)


def a05_render(lib, cols, rows, molData=None, meshData=None):
    """Same three flows with bounding boxes (A05/code.js:536-624): `initTrace` clips the primary ray to the
    (merged) bounds, each trace kernel first tests its own set's box."""
    bounds = _merged_bounds(molData, meshData)
    cam16 = mol_camera(bounds, cols, rows).toFloat32Array()
    pix = np.zeros((rows * cols, 4), dtype=np.uint8)
    rays = np.zeros(rows * cols, dtype=RAY)
    lib.a05_initTrace(pix, cam16, rays, H.bounds2AABB(bounds), cols, rows)
    if molData is not None:
        atoms, colors = pack_atoms(molData)
        lib.a05_molTrace(pix, cam16, rays, int(molData["size"]), atoms, colors, H.bounds2AABB(molData["bounds"]), cols, rows)
    if meshData is not None:
        t = prepare_a045_mesh(meshData)
        lib.a05_meshTrace(pix, cam16, rays, t["size"], t["pos"], t["normal"], t["index"], t["colors"], t["aabb"], cols, rows)
    return pix.reshape(rows, cols, 4), rays


def prepare_a06_mol(molData, n_slabs):
    """prepareMolTrace, A06/code.js:434-544."""
    atoms, colors, limits, idx = H.slabSplitMol(molData, n_slabs)
    return {"size": int(molData["size"]), "atoms": H.to_f32(atoms), "colors": H.to_f32(colors), "index": idx,
            "box": np.asarray(limits, dtype=np.uint32), "aabb": H.bounds2AABB(molData["bounds"]), "n": int(n_slabs)}


def prepare_a06_mesh(meshData, n_slabs):
    """prepareMeshTrace, A06/code.js:546-603."""
    pos, nor, idx, limits = H.slabSplitMesh(meshData, n_slabs)
    return {"size": int(meshData["nTriangles"]), "pos": H.to_f32(pos), "normal": H.to_f32(nor), "index": np.asarray(idx, dtype=np.uint32),
            "colors": H.to_f32(meshData["materials"]), "box": np.asarray(limits, dtype=np.uint32),
            "aabb": H.bounds2AABB(meshData["bounds"]), "n": int(n_slabs)}


def a06_render(lib, cols, rows, n_slabs=5, molData=None, meshData=None):
    """compute / computeTri / computeBoth of A06 (A06/code.js:632-729): 1-D slabs along x.
    Returns (pixels, rays, prep)."""
    bounds = _merged_bounds(molData, meshData)
    cam16 = mol_camera(bounds, cols, rows).toFloat32Array()
    pix = np.zeros((rows * cols, 4), dtype=np.uint8)
    rays = np.zeros(rows * cols, dtype=RAY)
    lib.a06_initTrace(pix, cam16, rays, H.bounds2AABB(bounds), cols, rows)
    prep = {"cam": cam16}
    if molData is not None:
        m = prep["mol"] = prepare_a06_mol(molData, n_slabs)
        lib.a06_molTrace(pix, cam16, rays, m["size"], m["atoms"], m["colors"], m["aabb"], m["n"], m["box"], cols, rows)
    if meshData is not None:
        t = prepare_a06_mesh(meshData, n_slabs)
        prep["mesh"] = t
        lib.a06_meshTrace(pix, cam16, rays, t["size"], t["pos"], t["normal"], t["index"], t["colors"], t["aabb"], t["n"], t["box"], cols, rows)
    return pix.reshape(rows, cols, 4), rays, prep


def prepare_a07_mol(molData, n_slabs):
    """prepareMolTrace, A07/code.js:434-483."""
    pos, idx, box = H.splitMolData(molData, n_slabs)
    return {"size": int(molData["size"]), "atoms": H.to_f32(pos), "index": np.asarray(idx, dtype=np.uint32),
            "colors": H.to_f32(molData["colorData"]), "box": np.asarray(box, dtype=np.uint32),
            "aabb": H.bounds2AABB(molData["bounds"]), "n": int(n_slabs)}


def prepare_a07_mesh(meshData, n_slabs):
    """prepareMeshTrace, A07/code.js:485-542."""
    pos, nor, box, idx = H.splitMeshData(meshData, n_slabs)
    return {"size": int(meshData["nTriangles"]), "pos": H.to_f32(pos), "normal": H.to_f32(nor),
            "index": np.asarray(idx, dtype=np.uint32), "colors": H.to_f32(meshData["materials"]),
            "box": np.asarray(box, dtype=np.uint32), "aabb": H.bounds2AABB(meshData["bounds"]), "n": int(n_slabs)}


def a07_render(lib, cols, rows, n_slabs=2, molData=None, meshData=None):
    """compute / computeTri / computeBoth, A07/code.js:571-668: `initTrace` against the (merged)
    bounds, then `molTrace` and/or `meshTrace` over the same ray buffer.  Returns
    (pixels, rays, prep)."""
    bounds = H.Bounds()
    if molData is not None and meshData is not None:
        bounds.merge(molData["bounds"])
        bounds.merge(meshData["bounds"])
    else:
        bounds = (molData or meshData)["bounds"]
    cam16 = mol_camera(bounds, cols, rows).toFloat32Array()
    pix = np.zeros((rows * cols, 4), dtype=np.uint8)
    rays = np.zeros(rows * cols, dtype=RAY)
    lib.a07_initTrace(pix, cam16, rays, H.bounds2AABB(bounds), cols, rows)
    prep = {"cam": cam16, "aabb": H.bounds2AABB(bounds)}
    if molData is not None:
        m = prep["mol"] = prepare_a07_mol(molData, n_slabs)
        lib.a07_molTrace(pix, cam16, rays, m["size"], m["atoms"], m["index"], m["colors"], m["aabb"], m["n"], m["box"], cols, rows)
    if meshData is not None:
        t = prep["mesh"] = prepare_a07_mesh(meshData, n_slabs)
        lib.a07_meshTrace(pix, cam16, rays, t["size"], t["pos"], t["normal"], t["index"], t["colors"], t["aabb"], t["n"], t["box"],
                          cols, rows)
    return pix.reshape(rows, cols, 4), rays, prep


def prepare_a089(scene, n_slabs=5):
    """prepareSphereTrace / prepareTriangleTrace of A08/A09 (A08/code.js:684-775)."""
    out = {"aabb": H.bounds2AABB(scene["bounds"]), "materials": H.splitMaterialData(scene), "n": int(n_slabs),
           "sphere": None, "triangle": None,
           "lights": [np.array([L.x, L.y, L.z, 1.0], dtype=np.float64).astype(np.float32) for L in scene["lights"]]}
    if len(scene["spheres"]) > 0:
        data, mat, box = H.splitSphereData(scene, n_slabs)
        out["sphere"] = {"data": H.to_f32(data), "matid": mat.astype(np.uint32), "box": box, "aabb": H.bounds2AABB(scene["sphereBounds"])}
    if len(scene["triangles"]) > 0:
        pos, nor, mat, box = H.splitTriangleData(scene, n_slabs)
        out["triangle"] = {"pos": H.to_f32(pos), "normal": H.to_f32(nor), "matid": mat.astype(np.uint32), "box": box,
                           "aabb": H.bounds2AABB(scene["triangleBounds"])}
    return out


def a08_render(lib, scene, cols, rows, n_slabs=5):
    """render(), A08/code.js:1194-1232: one ray per pixel, point lights, 2-D NDRange.  Quirk Q10:
    triangleShadowTrace is handed the SPHERE bounds (A08/code.js:918).  Returns (acu, pixels, state)."""
    prep = prepare_a089(scene, n_slabs)
    n = cols * rows
    st = {"rays": np.zeros(n, dtype=RAY), "pois": np.zeros(n, dtype=POI8), "shadow": np.zeros(n, dtype=RAY),
          "acu": np.zeros((n, 4), dtype=np.float32)}
    cam16 = scene["camera"].toFloat32Array()
    S, T, N = prep["sphere"], prep["triangle"], prep["n"]
    lib.a08_initTrace(st["acu"], st["rays"], st["pois"], prep["aabb"], cam16, cols, rows)
    if S:
        lib.a08_sphereTrace(cols, rows, st["pois"], st["rays"], S["data"], S["matid"], S["box"], S["aabb"], N)
    if T:
        lib.a08_triangleTrace(cols, rows, st["pois"], st["rays"], T["pos"], T["normal"], T["matid"], T["box"], T["aabb"], N)
    sphere_aabb = H.bounds2AABB(scene["sphereBounds"])
    for L in prep["lights"]:
        lib.a08_initShadowTrace(st["shadow"], st["pois"], cols, rows, L)
        if S:
            lib.a08_sphereShadowTrace(cols, rows, st["shadow"], S["data"], S["box"], S["aabb"], N)
        if T:
            lib.a08_triangleShadowTrace(cols, rows, st["shadow"], T["pos"], T["box"], sphere_aabb, N)
        lib.a08_sceneRender(st["acu"], st["pois"], st["shadow"], prep["materials"], n)
    pix = np.zeros((n, 4), dtype=np.uint8)
    lib.a08_copyToPixel(pix, st["acu"], float(np.float32(1.0 / len(prep["lights"]))), n)
    return st["acu"], pix.reshape(rows, cols, 4), st


def a09_render(lib, scene, cols, rows, rays_per_pixel, n_slabs=5, focal_length=None, lens_diameter=None):
    """render(), A09/code.js:1256-1294: thin-lens stratified primaries, 1-D NDRange over
    total_rays.  Returns (acu [total,4], pixels, state)."""
    prep = prepare_a089(scene, n_slabs)
    total = cols * rows * rays_per_pixel
    st = {"rays": np.zeros(total, dtype=RAY), "pois": np.zeros(total, dtype=POI8), "shadow": np.zeros(total, dtype=RAY),
          "acu": np.zeros((total, 4), dtype=np.float32)}
    cam16 = scene["camera"].toFloat32Array()
    fl = scene["focal_length"] if focal_length is None else focal_length
    ld = scene["lens_diameter"] if lens_diameter is None else lens_diameter
    S, T, N = prep["sphere"], prep["triangle"], prep["n"]
    lib.a09_initTrace(st["acu"], st["rays"], st["pois"], prep["aabb"], cam16, float(np.float32(fl)), float(np.float32(ld / 2.0)),
                      rays_per_pixel, cols, rows)
    if S:
        lib.a09_sphereTrace(total, st["pois"], st["rays"], S["data"], S["matid"], S["box"], S["aabb"], N)
    if T:
        lib.a09_triangleTrace(total, st["pois"], st["rays"], T["pos"], T["normal"], T["matid"], T["box"], T["aabb"], N)
    for L in prep["lights"]:
        lib.a09_initShadowTrace(st["shadow"], st["pois"], total, L)
        if S:
            lib.a09_sphereShadowTrace(total, st["shadow"], S["data"], S["box"], S["aabb"], N)
        if T:
            lib.a09_triangleShadowTrace(total, st["shadow"], T["pos"], T["box"], T["aabb"], N)
        lib.a09_sceneRender(st["acu"], st["pois"], st["shadow"], prep["materials"], total)
    pix = np.zeros((cols * rows, 4), dtype=np.uint8)
    lib.a09_copyToPixel(pix, st["acu"], float(np.float32(1.0 / (rays_per_pixel * len(prep["lights"])))), cols * rows, rays_per_pixel)
    return st["acu"], pix.reshape(rows, cols, 4), st
