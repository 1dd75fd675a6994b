"""Frame drivers of the earlier assignments (BASELINE.json configs 1-4) on the GPU -- host-side
mirrors of ``compute`` / ``computeTri`` / ``computeBoth`` / ``render`` of each ``code.js``:
same buffer preparation, same kernel sequence, one C-ABI launcher per reference kernel.

    A01  compute()                      A01/code.js:166-269
    A02  compute()  (one fused kernel)  A02/code.js:413-573
    A03  compute()  (initTrace+molTrace) A03/code.js:450-598
    A04  compute / computeTri / computeBoth   A04/code.js:520-606   (brute force, spheres + triangle soup)
    A05  compute / computeTri / computeBoth   A05/code.js:536-624   (+ bounding boxes)
    A06  compute / computeTri / computeBoth   A06/code.js:632-729   (1-D slabs along x)
    A07  compute / computeTri / computeBoth   A07/code.js:571-668
    A08  render()                       A08/code.js:1194-1232
    A09  render()                       A09/code.js:1256-1294

Every function takes a :class:`lib.Context`, leaves no device memory behind
(releaseCLResources) and returns host arrays.  ``timing=True`` additionally returns the device
time of the kernel sequence (CUDA events on the context's stream, uploads/readbacks excluded) --
the "grid-traversal ms/frame" metric of configs 1-4.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import host as H
from . import lib as L

RAY_BYTES = 48    # rt_struct_size("Ray", *)
POI8_BYTES = 48   # rt_struct_size("Poi", 8 | 9)


class _Frame:
    """Scoped device buffers + optional CUDA-event timing of the launches made through call()."""

    def __init__(self, ctx, timing=False):
        self.ctx, self.bufs, self.timing = ctx, [], timing
        self.ms = None
        if timing:
            import torch
            self._torch = torch
            self._stream = torch.cuda.ExternalStream(L.dll.rt_ctx_stream(ctx.h))
            self._e0, self._e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def alloc(self, nbytes):
        p = self.ctx.alloc(max(int(nbytes), 1))
        self.bufs.append(p)
        return p

    def upload(self, arr):
        p = self.ctx.upload(arr)
        self.bufs.append(p)
        return p

    def begin(self):
        if self.timing:
            self.ctx.finish()
            self._e0.record(self._stream)

    def end(self):
        if self.timing:
            self._e1.record(self._stream)
            self.ctx.finish()
            self.ms = self._e0.elapsed_time(self._e1)

    def close(self):
        for p in reversed(self.bufs):   # LIFO, like cl_resources (A10/code.js:1539-1546)
            self.ctx.free(p)
        self.bufs = []


def _ret(frame, *vals):
    return vals + (frame.ms,) if frame.timing else (vals if len(vals) > 1 else vals[0])


def camera_a01(cols, rows) -> np.ndarray:
    """A01/code.js:43-58,180-185: constants; rows before cols in the float16."""
    return np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 2.66, 2.0, rows, cols], dtype=np.float32)


def a01_compute(ctx, cols, rows, timing=False):
    f = _Frame(ctx, timing)
    try:
        d_pix = f.alloc(4 * cols * rows)
        cam = camera_a01(cols, rows)
        f.begin()
        ctx.call("rt_a01_raytrace", d_pix, L.hptr(cam))
        f.end()
        pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
    finally:
        f.close()
    return _ret(f, pix)


def pack_atoms(molData):
    """A03/code.js:524-537: Float32Array(size*4) of (x,y,z,radius) and of the element colour;
    records past the end of atomData read `undefined` -> NaN (quirk Q13)."""
    n = int(molData["size"])
    ad = np.asarray(molData["atomData"], dtype=np.float64).reshape(-1, 4)
    rd = np.asarray(molData["radiusData"], dtype=np.float64)
    cd = np.asarray(molData["colorData"], dtype=np.float64).reshape(-1, 4)
    atoms, colors = np.full((n, 4), np.nan), np.full((n, 4), np.nan)
    m = min(n, len(ad))
    ids = ad[:m, 0].astype(np.int64)
    atoms[:m, :3] = ad[:m, 1:4]
    atoms[:m, 3] = rd[ids]
    colors[:m] = cd[ids]
    return atoms.astype(np.float32).reshape(-1), colors.astype(np.float32).reshape(-1)


def _mol_camera(bounds, cols, rows):
    cam = H.Camera()
    cam.defaultInit()
    cam.set(bounds, cols, rows)
    return cam.toFloat32Array()


def a02_compute(ctx, molData, cols, rows, timing=False):
    f = _Frame(ctx, timing)
    try:
        atoms, colors = pack_atoms(molData)
        d_atoms, d_colors, d_pix = f.upload(atoms), f.upload(colors), f.alloc(4 * cols * rows)
        cam = _mol_camera(molData["bounds"], cols, rows)
        f.begin()
        ctx.call("rt_a02_raytrace", d_pix, L.hptr(cam), int(molData["size"]), d_atoms, d_colors)
        f.end()
        pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
    finally:
        f.close()
    return _ret(f, pix)


def a03_compute(ctx, molData, cols, rows, timing=False):
    """Returns (pixels, ray mint per pixel)."""
    f = _Frame(ctx, timing)
    try:
        atoms, colors = pack_atoms(molData)
        d_atoms, d_colors = f.upload(atoms), f.upload(colors)
        d_pix, d_rays = f.alloc(4 * cols * rows), f.alloc(RAY_BYTES * cols * rows)
        cam = _mol_camera(molData["bounds"], cols, rows)
        f.begin()
        ctx.call("rt_a03_initTrace", d_pix, L.hptr(cam), d_rays)
        ctx.call("rt_a03_molTrace", d_pix, L.hptr(cam), d_rays, int(molData["size"]), d_atoms, d_colors)
        f.end()
        pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
        rays = ctx.download(d_rays, np.float32, 12 * cols * rows).reshape(-1, 12)
    finally:
        f.close()
    return _ret(f, pix, rays[:, 8].copy())


def _both_bounds(molData, meshData, who):
    if molData is None and meshData is None:
        raise ValueError("%s: need molData and/or meshData" % who)
    if molData is not None and meshData is not None:
        bounds = H.Bounds()
        bounds.merge(molData["bounds"])
        bounds.merge(meshData["bounds"])
        return bounds
    return (molData or meshData)["bounds"]


def _mesh_soup(f, meshData):
    """prepareMeshTrace of A04/A05 (A04/code.js:448-492)."""
    return (f.upload(H.toPosArray(meshData)), f.upload(H.toNormalArray(meshData)),
            f.upload(np.ascontiguousarray(meshData["materialIndices"], dtype=np.uint32)),
            f.upload(np.asarray(meshData["materials"], dtype=np.float64).astype(np.float32)))


def a045_compute(ctx, cols, rows, molData=None, meshData=None, boxes=False, timing=False):
    """compute / computeTri / computeBoth of A04 (``boxes=False``) or A05 (``boxes=True``): initTrace, then the
    brute-force molTrace and/or meshTrace over one ray buffer.  Returns (pixels, ray maxt per pixel[, ms])."""
    bounds = _both_bounds(molData, meshData, "a045_compute")
    a = "rt_a05_" if boxes else "rt_a04_"
    f = _Frame(ctx, timing)
    try:
        cam = _mol_camera(bounds, cols, rows)
        d_pix, d_rays = f.alloc(4 * cols * rows), f.alloc(RAY_BYTES * cols * rows)
        mol = mesh = None
        if molData is not None:
            atoms, colors = pack_atoms(molData)
            mol = (f.upload(atoms), f.upload(colors))
        if meshData is not None:
            mesh = _mesh_soup(f, meshData)
        f.begin()
        if boxes:
            ctx.call(a + "initTrace", d_pix, L.hptr(cam), d_rays, L.hptr(H.bounds2AABB(bounds)))
        else:
            ctx.call(a + "initTrace", d_pix, L.hptr(cam), d_rays)
        if mol:
            extra = (L.hptr(H.bounds2AABB(molData["bounds"])),) if boxes else ()
            ctx.call(a + "molTrace", d_pix, L.hptr(cam), d_rays, int(molData["size"]), mol[0], mol[1], *extra)
        if mesh:
            extra = (L.hptr(H.bounds2AABB(meshData["bounds"])),) if boxes else ()
            ctx.call(a + "meshTrace", d_pix, L.hptr(cam), d_rays, int(meshData["nTriangles"]), mesh[0], mesh[1], mesh[2], mesh[3], *extra)
        f.end()
        pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
        rays = ctx.download(d_rays, np.float32, 12 * cols * rows).reshape(-1, 12)
    finally:
        f.close()
    return _ret(f, pix, rays[:, 9].copy())


def a04_compute(ctx, cols, rows, molData=None, meshData=None, timing=False):
    return a045_compute(ctx, cols, rows, molData, meshData, False, timing)


def a05_compute(ctx, cols, rows, molData=None, meshData=None, timing=False):
    return a045_compute(ctx, cols, rows, molData, meshData, True, timing)


def a04_raytrace(ctx, molData, cols, rows):
    """The fused kernel Assignment 4 still carries (A04/code.cl:317-364); its code.js no longer launches it."""
    f = _Frame(ctx)
    try:
        atoms, colors = pack_atoms(molData)
        d_atoms, d_colors, d_pix = f.upload(atoms), f.upload(colors), f.alloc(4 * cols * rows)
        cam = _mol_camera(molData["bounds"], cols, rows)
        ctx.call("rt_a04_raytrace", d_pix, L.hptr(cam), int(molData["size"]), d_atoms, d_colors)
        pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
    finally:
        f.close()
    return pix


def a06_compute(ctx, cols, rows, n_slabs=5, molData=None, meshData=None, timing=False):
    """compute / computeTri / computeBoth of A06: x slabs built on the GPU (rt_slab_build_*), initTrace against the
    (merged) bounds, then molTrace and/or meshTrace.  Returns (pixels, ray maxt per pixel[, ms])."""
    bounds = _both_bounds(molData, meshData, "a06_compute")
    f = _Frame(ctx, timing)
    grids = []
    try:
        cam = _mol_camera(bounds, cols, rows)
        d_pix, d_rays = f.alloc(4 * cols * rows), f.alloc(RAY_BYTES * cols * rows)
        mol = mesh = None
        if molData is not None:
            g = H.slabSplitMolData(ctx, molData, n_slabs)
            grids.append(g)
            # colorData per slab reference (A06/code.js:511-514) = element colour looked up through the reference's element index
            ids = ctx.download(g.matid, np.uint32, g.n_refs) if g.n_refs else np.zeros(0, np.uint32)
            col = np.asarray(molData["colorData"], dtype=np.float64).reshape(-1, 4)[ids].astype(np.float32).reshape(-1)
            mol = (g, f.upload(col) if len(col) else f.alloc(16))
        if meshData is not None:
            g = H.slabSplitMeshData(ctx, meshData, n_slabs)
            grids.append(g)
            mesh = (g, f.upload(np.asarray(meshData["materials"], dtype=np.float64).astype(np.float32)))
        f.begin()
        ctx.call("rt_a06_initTrace", d_pix, L.hptr(cam), d_rays, L.hptr(H.bounds2AABB(bounds)))
        if mol:
            g, d_col = mol
            ctx.call("rt_a06_molTrace", d_pix, L.hptr(cam), d_rays, int(molData["size"]), g.prim, d_col, L.hptr(H.bounds2AABB(molData["bounds"])),
                     int(n_slabs), g.box_size)
        if mesh:
            g, d_col = mesh
            ctx.call("rt_a06_meshTrace", d_pix, L.hptr(cam), d_rays, int(meshData["nTriangles"]), g.prim, g.normal, g.matid, d_col,
                     L.hptr(H.bounds2AABB(meshData["bounds"])), int(n_slabs), g.box_size)
        f.end()
        pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
        rays = ctx.download(d_rays, np.float32, 12 * cols * rows).reshape(-1, 12)
    finally:
        for g in grids:
            L.dll.rt_grid_release(ctx.h, C.byref(g))
        f.close()
    return _ret(f, pix, rays[:, 9].copy())


def a07_compute(ctx, cols, rows, n_slabs=2, molData=None, meshData=None, timing=False):
    """compute (molData), computeTri (meshData) or computeBoth (both): initTrace against the
    (merged) bounds, then molTrace and/or meshTrace over one ray buffer.  Returns
    (pixels, ray maxt per pixel[, ms])."""
    if molData is None and meshData is None:
        raise ValueError("a07_compute: need molData and/or meshData")
    if molData is not None and meshData is not None:
        bounds = H.Bounds()
        bounds.merge(molData["bounds"])
        bounds.merge(meshData["bounds"])
    else:
        bounds = (molData or meshData)["bounds"]
    f = _Frame(ctx, timing)
    grids = []
    try:
        cam = _mol_camera(bounds, cols, rows)
        aabb = H.bounds2AABB(bounds)
        d_pix, d_rays = f.alloc(4 * cols * rows), f.alloc(RAY_BYTES * cols * rows)
        mol = mesh = None
        if molData is not None:
            g = H.splitMolData(ctx, molData, n_slabs)
            grids.append(g)
            mol = (g, f.upload(np.asarray(molData["colorData"], dtype=np.float64).astype(np.float32)), H.bounds2AABB(molData["bounds"]))
        if meshData is not None:
            g = H.splitMeshData(ctx, meshData, n_slabs)
            grids.append(g)
            mesh = (g, f.upload(np.asarray(meshData["materials"], dtype=np.float64).astype(np.float32)), H.bounds2AABB(meshData["bounds"]))
        f.begin()
        ctx.call("rt_a07_initTrace", d_pix, L.hptr(cam), d_rays, L.hptr(aabb))
        if mol:
            g, d_col, bb = mol
            ctx.call("rt_a07_molTrace", d_pix, L.hptr(cam), d_rays, int(molData["size"]), g.prim, g.matid, d_col, L.hptr(bb), int(n_slabs),
                     g.box_size)
        if mesh:
            g, d_col, bb = mesh
            ctx.call("rt_a07_meshTrace", d_pix, L.hptr(cam), d_rays, int(meshData["nTriangles"]), g.prim, g.normal, g.matid, d_col, L.hptr(bb),
                     int(n_slabs), g.box_size)
        f.end()
        pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
        rays = ctx.download(d_rays, np.float32, 12 * cols * rows).reshape(-1, 12)
    finally:
        for g in grids:
            L.dll.rt_grid_release(ctx.h, C.byref(g))
        f.close()
    return _ret(f, pix, rays[:, 9].copy())


def _a089_sets(ctx, f, scene, n_slabs, grids):
    S = T = None
    if len(scene["spheres"]) > 0:
        g = H.splitSphereData(ctx, scene, n_slabs)
        grids.append(g)
        S = (g, H.bounds2AABB(scene["sphereBounds"]))
    if len(scene["triangles"]) > 0:
        g = H.splitTriangleData(ctx, scene, n_slabs)
        grids.append(g)
        T = (g, H.bounds2AABB(scene["triangleBounds"]))
    return S, T


def _light_pos(v):
    return np.array([v.x, v.y, v.z, 1.0], dtype=np.float64).astype(np.float32)


def a08_render(ctx, scene, cols, rows, n_slabs=5, timing=False):
    """render() of A08: one ray per pixel, point lights.  Quirk Q10: triangleShadowTrace is handed
    the SPHERE bounds (A08/code.js:918).  Returns (acu [pixels,4], pixels, hit matId, ray maxt[, ms])."""
    n = cols * rows
    f = _Frame(ctx, timing)
    grids = []
    try:
        S, T = _a089_sets(ctx, f, scene, n_slabs, grids)
        d_acu, d_rays, d_pois, d_sh = f.alloc(16 * n), f.alloc(RAY_BYTES * n), f.alloc(POI8_BYTES * n), f.alloc(RAY_BYTES * n)
        d_mat, d_pix = f.upload(H.splitMaterialData(scene)), f.alloc(4 * n)
        cam, aabb, sph_bb = scene["camera"].toFloat32Array(), H.bounds2AABB(scene["bounds"]), H.bounds2AABB(scene["sphereBounds"])
        lights = [_light_pos(v) for v in scene["lights"]]
        f.begin()
        ctx.call("rt_a08_initTrace", d_acu, d_rays, d_pois, L.hptr(aabb), L.hptr(cam))
        if S:
            ctx.call("rt_a08_sphereTrace", cols, rows, d_pois, d_rays, S[0].prim, S[0].matid, S[0].box_size, L.hptr(S[1]), int(n_slabs))
        if T:
            ctx.call("rt_a08_triangleTrace", cols, rows, d_pois, d_rays, T[0].prim, T[0].normal, T[0].matid, T[0].box_size, L.hptr(T[1]), int(n_slabs))
        for lp in lights:
            ctx.call("rt_a08_initShadowTrace", d_sh, d_pois, cols, rows, L.hptr(lp))
            if S:
                ctx.call("rt_a08_sphereShadowTrace", cols, rows, d_sh, S[0].prim, S[0].box_size, L.hptr(S[1]), int(n_slabs))
            if T:
                ctx.call("rt_a08_triangleShadowTrace", cols, rows, d_sh, T[0].prim, T[0].box_size, L.hptr(sph_bb), int(n_slabs))
            ctx.call("rt_a08_sceneRender", d_acu, d_pois, d_sh, d_mat, n)
        ctx.call("rt_a08_copyToPixel", d_pix, d_acu, float(np.float32(1.0 / len(lights))), n)
        f.end()
        out = _a089_readback(ctx, d_acu, d_pix, d_pois, d_rays, n, cols, rows)
    finally:
        for g in grids:
            L.dll.rt_grid_release(ctx.h, C.byref(g))
        f.close()
    return _ret(f, *out)


def a09_render(ctx, scene, cols, rows, rays_per_pixel, n_slabs=5, focal_length=None, lens_diameter=None, timing=False):
    """render() of A09: thin-lens stratified primaries (rays_per_pixel a perfect square), 1-D
    launches over total_rays.  Returns (acu [total,4], pixels, hit matId, ray maxt[, ms])."""
    total = cols * rows * rays_per_pixel
    fl = scene["focal_length"] if focal_length is None else focal_length
    ld = scene["lens_diameter"] if lens_diameter is None else lens_diameter
    f = _Frame(ctx, timing)
    grids = []
    try:
        S, T = _a089_sets(ctx, f, scene, n_slabs, grids)
        d_acu, d_rays, d_pois, d_sh = f.alloc(16 * total), f.alloc(RAY_BYTES * total), f.alloc(POI8_BYTES * total), f.alloc(RAY_BYTES * total)
        d_mat, d_pix = f.upload(H.splitMaterialData(scene)), f.alloc(4 * cols * rows)
        cam, aabb = scene["camera"].toFloat32Array(), H.bounds2AABB(scene["bounds"])
        lights = [_light_pos(v) for v in scene["lights"]]
        f.begin()
        ctx.call("rt_a09_initTrace", d_acu, d_rays, d_pois, L.hptr(aabb), L.hptr(cam), float(np.float32(fl)), float(np.float32(ld / 2.0)),
                 int(rays_per_pixel))
        if S:
            ctx.call("rt_a09_sphereTrace", total, d_pois, d_rays, S[0].prim, S[0].matid, S[0].box_size, L.hptr(S[1]), int(n_slabs))
        if T:
            ctx.call("rt_a09_triangleTrace", total, d_pois, d_rays, T[0].prim, T[0].normal, T[0].matid, T[0].box_size, L.hptr(T[1]), int(n_slabs))
        for lp in lights:
            ctx.call("rt_a09_initShadowTrace", d_sh, d_pois, total, L.hptr(lp))
            if S:
                ctx.call("rt_a09_sphereShadowTrace", total, d_sh, S[0].prim, S[0].box_size, L.hptr(S[1]), int(n_slabs))
            if T:
                ctx.call("rt_a09_triangleShadowTrace", total, d_sh, T[0].prim, T[0].box_size, L.hptr(T[1]), int(n_slabs))
            ctx.call("rt_a09_sceneRender", d_acu, d_pois, d_sh, d_mat, total)
        ctx.call("rt_a09_copyToPixel", d_pix, d_acu, float(np.float32(1.0 / (rays_per_pixel * len(lights)))), cols * rows, int(rays_per_pixel))
        f.end()
        out = _a089_readback(ctx, d_acu, d_pix, d_pois, d_rays, total, cols, rows)
    finally:
        for g in grids:
            L.dll.rt_grid_release(ctx.h, C.byref(g))
        f.close()
    return _ret(f, *out)


def a089_render_fused(ctx, scene, cols, rows, assignment, rays_per_pixel=1, n_slabs=5, focal_length=None, lens_diameter=None, timing=False):
    """render() of Assignment 8 (``assignment=8``: pinhole, one ray per pixel) or 9 (thin lens, ``rays_per_pixel`` a perfect square)
    through rt_a089_render_frame: the whole launcher sequence of :func:`a08_render` / :func:`a09_render` in one kernel plus
    copyToPixel.  Same return values, bit for bit."""
    rpp = 1 if assignment == 8 else int(rays_per_pixel)
    total = cols * rows * rpp
    f = _Frame(ctx, timing)
    grids = []
    try:
        S, T = _a089_sets(ctx, f, scene, n_slabs, grids)
        d_acu, d_mat, d_pix = f.alloc(16 * total), f.upload(H.splitMaterialData(scene)), f.alloc(4 * cols * rows)
        d_id, d_maxt = f.alloc(4 * total), f.alloc(4 * total)
        lights = np.concatenate([_light_pos(v) for v in scene["lights"]]) if scene["lights"] else np.zeros(4, np.float32)
        fr = L.A089Frame()
        if S:
            fr.spheres, fr.s_matid, fr.s_box_size, fr.s_n_slabs = S[0].prim, S[0].matid, S[0].box_size, int(n_slabs)
            fr.s_bound[:] = [float(v) for v in S[1]]
        if T:
            fr.t_pos, fr.t_normal, fr.t_matid, fr.t_box_size, fr.t_n_slabs = T[0].prim, T[0].normal, T[0].matid, T[0].box_size, int(n_slabs)
            fr.t_bound[:] = [float(v) for v in T[1]]
            shadow_bb = H.bounds2AABB(scene["sphereBounds"]) if assignment == 8 else T[1]   # quirk Q10
            fr.t_shadow_bound[:] = [float(v) for v in shadow_bb]
        fr.material = d_mat
        fr.light_pos, fr.n_lights = L.hptr(lights), len(scene["lights"])
        fr.bound[:] = [float(v) for v in H.bounds2AABB(scene["bounds"])]
        fr.fcam[:] = [float(v) for v in scene["camera"].toFloat32Array()]
        if assignment != 8:
            fl = scene["focal_length"] if focal_length is None else focal_length
            ld = scene["lens_diameter"] if lens_diameter is None else lens_diameter
            fr.focal_length, fr.lens_rad, fr.thin_lens = float(np.float32(fl)), float(np.float32(ld / 2.0)), 1
        fr.rays_per_pixel = rpp
        f.begin()
        ctx.check(L.dll.rt_a089_render_frame(ctx.h, C.byref(fr), d_acu, d_id, d_maxt))
        if assignment == 8:
            ctx.call("rt_a08_copyToPixel", d_pix, d_acu, float(np.float32(1.0 / len(scene["lights"]))), cols * rows)
        else:
            ctx.call("rt_a09_copyToPixel", d_pix, d_acu, float(np.float32(1.0 / (rpp * len(scene["lights"])))), cols * rows, rpp)
        f.end()
        acu = ctx.download(d_acu, np.float32, 4 * total).reshape(-1, 4)
        pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
        out = (acu, pix, ctx.download(d_id, np.int32, total), ctx.download(d_maxt, np.float32, total))
    finally:
        for g in grids:
            L.dll.rt_grid_release(ctx.h, C.byref(g))
        f.close()
    return _ret(f, *out)


def _a089_readback(ctx, d_acu, d_pix, d_pois, d_rays, total, cols, rows):
    acu = ctx.download(d_acu, np.float32, 4 * total).reshape(-1, 4)
    pix = ctx.download(d_pix, np.uint8, 4 * cols * rows).reshape(rows, cols, 4)
    pois = ctx.download(d_pois, np.int32, 12 * total).reshape(-1, 12)
    rays = ctx.download(d_rays, np.float32, 12 * total).reshape(-1, 12)
    return acu, pix, pois[:, 8].copy(), rays[:, 9].copy()
