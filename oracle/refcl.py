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
_SIZEOF = ["a03_sizeofRay", "a07_sizeofRay", "a08_sizeofRay", "a08_sizeofPoi", "a09_sizeofRay", "a09_sizeofPoi",
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
            fn = getattr(self._dll, prefix + name)
            fn.argtypes = sig
            fn.restype = None
            setattr(self, name, self._wrap(fn))
        for name in _SIZEOF:
            fn = getattr(self._dll, prefix + name)
            fn.argtypes = []
            fn.restype = _U
            setattr(self, name, fn)
        self.num_threads = getattr(self._dll, prefix + "num_threads")
        self.num_threads.restype = _I
        self._set_stats = getattr(self._dll, prefix + "set_stats", None)
        if self._set_stats is not None:
            self._set_stats.argtypes = [_P, _P, _P]
        self._is_instr = getattr(self._dll, prefix + "is_instrumented", None)

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
