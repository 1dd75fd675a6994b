"""ctypes binding of librt2015.so -- exactly the entry points ``include/rt2015.h`` declares.

This is the Python twin of the N-API addon sketched in INTEGRATION.md: TypedArray/numpy
buffers in, opaque device pointers and status codes out.  No compute happens here and there
is no fallback: a missing library is an ImportError, a missing GPU is a RuntimeError from
``Context()``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT2015_LIB") or os.path.join(_HERE, "librt2015.so")   # override: A/B builds of the same ABI
if not os.path.exists(LIB_PATH):
    raise ImportError(
        "librt2015.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` or "
        "`make -C 2015-raytracing_b200/csrc` (there is no CPU fallback)")
dll = C.CDLL(LIB_PATH)

P, U, F, I, Z = C.c_void_p, C.c_uint, C.c_float, C.c_int, C.c_size_t
PP = C.POINTER(C.c_void_p)
ULL = C.c_ulonglong


class Grid(C.Structure):
    """rt_grid (include/rt2015.h)."""
    _fields_ = [("prim", P), ("normal", P), ("matid", P), ("box_size", P), ("occupancy", P), ("n_refs", U), ("n_slabs", U),
                ("kind", U), ("_reserved", U)]


class MeshXform(C.Structure):
    """rt_mesh_xform."""
    _fields_ = [("do_normalize", I), ("center", C.c_double * 3), ("maxdim", C.c_double), ("scale", C.c_double * 3),
                ("translate", C.c_double * 3)]


class MeshData(C.Structure):
    """rt_mesh_data."""
    _fields_ = [("n_triangles", C.c_uint), ("n_materials", C.c_uint), ("positions", C.POINTER(C.c_double)), ("normals", C.POINTER(C.c_double)),
                ("material_indices", C.POINTER(C.c_uint)), ("materials", C.POINTER(C.c_double)), ("bounds_min", C.c_double * 3),
                ("bounds_max", C.c_double * 3)]


class MolData(C.Structure):
    """rt_mol_data."""
    _fields_ = [("size", C.c_uint), ("n_records", C.c_uint), ("n_elements", C.c_uint), ("atom_data", C.POINTER(C.c_double)),
                ("color_data", C.POINTER(C.c_double)), ("radius_data", C.POINTER(C.c_double)), ("bounds_min", C.c_double * 3),
                ("bounds_max", C.c_double * 3)]


class A089Frame(C.Structure):
    """rt_a089_frame."""
    _fields_ = [("spheres", P), ("s_matid", P), ("s_box_size", P), ("s_bound", F * 8), ("s_n_slabs", U),
                ("t_pos", P), ("t_normal", P), ("t_matid", P), ("t_box_size", P), ("t_bound", F * 8), ("t_n_slabs", U),
                ("t_shadow_bound", F * 8), ("material", P), ("light_pos", P), ("n_lights", U), ("bound", F * 8), ("fcam", F * 16),
                ("focal_length", F), ("lens_rad", F), ("rays_per_pixel", U), ("thin_lens", U)]


class RenderOpts(C.Structure):
    """rt_render_opts."""
    _fields_ = [("cols", U), ("rows", U), ("rays_per_pixel", U), ("depth", U), ("focal_length", F), ("lens_rad", F),
                ("slot_begin", U), ("slot_count", U), ("mode", U), ("tile_slots", U)]


_SIGS = {
    "rt_ctx_create": ([I, PP], I),
    "rt_ctx_destroy": ([P], I),
    "rt_last_error_string": ([P], C.c_char_p),
    "rt_finish": ([P], I),
    "rt_ctx_stream": ([P], P),
    "rt_device_info": ([P, C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(Z), C.POINTER(Z)], I),
    "rt_buffer_create": ([P, Z, PP], I),
    "rt_buffer_release": ([P, P], I),
    "rt_buffer_write": ([P, P, Z, Z, P], I),
    "rt_buffer_read": ([P, P, Z, Z, P], I),
    "rt_buffer_fill": ([P, P, I, Z], I),
    "rt_struct_size": ([C.c_char_p, I], U),
    "rt_set_walk_stats": ([P, P, P, P], I),
    "rt_set_walk_totals": ([P, P], I),
    "rt_a10_initAcu": ([P, P, U], I),
    "rt_a10_initTrace": ([P, P, P, P, P, P, F, F, U], I),
    "rt_a10_bouncePaths": ([P, P, P, P, U], I),
    "rt_a10_lightRender": ([P, P, P, P, P, U], I),
    "rt_a10_initShadowTrace": ([P, P, P, U, P, P], I),
    "rt_a10_sphereTrace": ([P, U, P, P, P, P, P, P, U], I),
    "rt_a10_triangleTrace": ([P, U, P, P, P, P, P, P, P, U], I),
    "rt_a10_meshTrace": ([P, U, P, P, P, P, P, U, P, U], I),
    "rt_a10_sphereShadowTrace": ([P, U, P, P, P, P, U], I),
    "rt_a10_triangleShadowTrace": ([P, U, P, P, P, P, U], I),
    "rt_a10_sceneRender": ([P, P, P, P, P, P, U], I),
    "rt_a10_copyToPixel": ([P, P, P, F, U, U], I),
    "rt_a01_raytrace": ([P, P, P], I),
    "rt_a02_raytrace": ([P, P, P, U, P, P], I),
    "rt_a03_initTrace": ([P, P, P, P], I),
    "rt_a03_molTrace": ([P, P, P, P, U, P, P], I),
    "rt_a04_initTrace": ([P, P, P, P], I),
    "rt_a04_molTrace": ([P, P, P, P, U, P, P], I),
    "rt_a04_meshTrace": ([P, P, P, P, U, P, P, P, P], I),
    "rt_a04_raytrace": ([P, P, P, U, P, P], I),
    "rt_a05_initTrace": ([P, P, P, P, P], I),
    "rt_a05_molTrace": ([P, P, P, P, U, P, P, P], I),
    "rt_a05_meshTrace": ([P, P, P, P, U, P, P, P, P, P], I),
    "rt_a06_initTrace": ([P, P, P, P, P], I),
    "rt_a06_molTrace": ([P, P, P, P, U, P, P, P, U, P], I),
    "rt_a06_meshTrace": ([P, P, P, P, U, P, P, P, P, P, U, P], I),
    "rt_a07_initTrace": ([P, P, P, P, P], I),
    "rt_a07_molTrace": ([P, P, P, P, U, P, P, P, P, U, P], I),
    "rt_a07_meshTrace": ([P, P, P, P, U, P, P, P, P, P, U, P], I),
    "rt_a08_initTrace": ([P, P, P, P, P, P], I),
    "rt_a08_initShadowTrace": ([P, P, P, U, U, P], I),
    "rt_a08_sphereTrace": ([P, U, U, P, P, P, P, P, P, U], I),
    "rt_a08_triangleTrace": ([P, U, U, P, P, P, P, P, P, P, U], I),
    "rt_a08_sphereShadowTrace": ([P, U, U, P, P, P, P, U], I),
    "rt_a08_triangleShadowTrace": ([P, U, U, P, P, P, P, U], I),
    "rt_a08_sceneRender": ([P, P, P, P, P, U], I),
    "rt_a08_copyToPixel": ([P, P, P, F, U], I),
    "rt_a09_initTrace": ([P, P, P, P, P, P, F, F, U], I),
    "rt_a09_initShadowTrace": ([P, P, P, U, P], I),
    "rt_a09_sphereTrace": ([P, U, P, P, P, P, P, P, U], I),
    "rt_a09_triangleTrace": ([P, U, P, P, P, P, P, P, P, U], I),
    "rt_a09_sphereShadowTrace": ([P, U, P, P, P, P, U], I),
    "rt_a09_triangleShadowTrace": ([P, U, P, P, P, P, U], I),
    "rt_a09_sceneRender": ([P, P, P, P, P, U], I),
    "rt_a09_copyToPixel": ([P, P, P, F, U, U], I),
    "rt_a089_render_frame": ([P, C.POINTER(A089Frame), P, P, P], I),
    "rt_parse_mesh_json": ([C.c_char_p, Z, C.POINTER(MeshData)], I),
    "rt_mesh_data_free": ([C.POINTER(MeshData)], None),
    "rt_parse_pdb": ([C.c_char_p, Z, C.POINTER(MolData)], I),
    "rt_mol_data_free": ([C.POINTER(MolData)], None),
    "rt_grid_build_spheres": ([P, P, P, U, P, P, U, C.POINTER(Grid)], I),
    "rt_grid_build_triangles": ([P, P, P, P, U, P, P, U, C.POINTER(MeshXform), C.POINTER(Grid)], I),
    "rt_slab_build_spheres": ([P, P, P, U, C.c_double, C.c_double, U, C.POINTER(Grid)], I),
    "rt_slab_build_triangles": ([P, P, P, P, U, C.c_double, C.c_double, U, C.POINTER(Grid)], I),
    "rt_grid_release": ([P, C.POINTER(Grid)], I),
    "rt_scene_create": ([P, PP], I),
    "rt_scene_destroy": ([P], I),
    "rt_scene_set_bounds": ([P, P], I),
    "rt_scene_set_materials": ([P, P, U], I),
    "rt_scene_add_set": ([P, C.POINTER(Grid), P, I, U], I),
    "rt_scene_add_light": ([P, P, P, P], I),
    "rt_scene_probe_empty_walks": ([P, U, P, U, P], I),
    "rt_render_create": ([P, P, C.POINTER(RenderOpts), PP], I),
    "rt_render_destroy": ([P], I),
    "rt_render_set_seeds": ([P, P, Z, I], I),
    "rt_render_execute": ([P, P, P], I),
    "rt_render_accum_image": ([P, PP], I),
    "rt_render_read_accum": ([P, P], I),
    "rt_render_read_seeds": ([P, P, Z], I),
    "rt_render_write_local_seeds": ([P, P, Z], I),
    "rt_render_write_local_seeds_async": ([P, P, Z], I),
    "rt_render_export_state": ([P, P, P, C.POINTER(U)], I),
    "rt_render_import_state": ([P, P, P, U], I),
    "rt_render_set_profile": ([P, I], I),
    "rt_render_read_profile": ([P, C.POINTER(ULL * 16)], I),
    "rt_render_read_profile_sets": ([P, C.POINTER(ULL * 128)], I),
    "rt_render_set_timing": ([P, I], I),
    "rt_render_read_timing": ([P, C.POINTER(F * 8), C.POINTER(U * 8)], I),
    "rt_accum_to_pixel": ([P, P, P, F, U], I),
    "rt_render_stats": ([P, C.POINTER(ULL), C.POINTER(ULL), C.POINTER(U), C.POINTER(F)], I),
    "rt_comm_unique_id": ([P], I),
    "rt_comm_create": ([P, I, I, P, PP], I),
    "rt_comm_destroy": ([P], I),
    "rt_render_reduce": ([P, P, I], I),
}
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(dll, _name)
    _fn.argtypes = _args
    _fn.restype = _res

EXPORTS = sorted(_SIGS)


class RtError(RuntimeError):
    pass


def hptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class Context:
    """rt_ctx -- one per GPU; replaces the WebCL context + in-order command queue
    (Assign10-Path_Tracing/code.js:576-608)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = dll.rt_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise RtError("rt_ctx_create(device=%d) failed with %d%s" % (
                device, rc, " (no CUDA device; there is no CPU fallback)" if rc == -4 else ""))
        self.h = h
        self._buffers = []

    def check(self, rc):
        if rc != 0:
            raise RtError("librt2015 error %d: %s" % (rc, dll.rt_last_error_string(self.h).decode()))

    def call(self, name, *args):
        self.check(getattr(dll, name)(self.h, *args))

    # -- buffers (ctx.createBuffer / enqueueWriteBuffer / enqueueReadBuffer) --
    def alloc(self, nbytes) -> int:
        p = C.c_void_p()
        self.check(dll.rt_buffer_create(self.h, int(nbytes), C.byref(p)))
        self._buffers.append(p.value)
        return p.value

    def upload(self, arr) -> int:
        arr = np.ascontiguousarray(arr)
        p = self.alloc(max(arr.nbytes, 1))
        if arr.nbytes:
            self.check(dll.rt_buffer_write(self.h, p, 0, arr.nbytes, arr.ctypes.data))
        return p

    def write(self, dptr, arr, offset=0):
        arr = np.ascontiguousarray(arr)
        self.check(dll.rt_buffer_write(self.h, dptr, offset, arr.nbytes, arr.ctypes.data))

    def download(self, dptr, dtype, count, offset=0) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            self.check(dll.rt_buffer_read(self.h, dptr, offset, out.nbytes, out.ctypes.data))
        return out

    def free(self, dptr):
        if dptr in self._buffers:
            self._buffers.remove(dptr)
        self.check(dll.rt_buffer_release(self.h, dptr))

    def finish(self):
        self.check(dll.rt_finish(self.h))

    def device_info(self):
        sm, ma, mi, l2, mem = I(), I(), I(), Z(), Z()
        self.check(dll.rt_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(l2), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "l2_bytes": l2.value, "total_mem": mem.value}

    def close(self):
        """releaseCLResources (A10/code.js:1539-1552): LIFO release, then the context."""
        if self.h:
            while self._buffers:
                dll.rt_buffer_release(self.h, self._buffers.pop())
            dll.rt_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
