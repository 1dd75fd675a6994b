"""2015-raytracing_b200 -- B200-native (sm_100a CUDA) implementation of the data-parallel hot
path of eaymerich/2015-RayTracing behind the reference's own kernel/host interface.

The compute lives in ``librt2015.so`` (C ABI declared in ``include/rt2015.h``); this package
is the host-side mirror of the reference's ``code.js`` surface (loaders, camera and light
packing, split*Data, preRender/executeRender/postRender) on top of that ABI.  There is no
CPU fallback: importing :mod:`.lib` raises if the CUDA library has not been built, and
creating a context raises if no CUDA device is present.

The directory name is not a valid Python identifier; import it with
``importlib.import_module("2015-raytracing_b200")`` (see ``__graft_entry__.py``).
"""
from . import lib  # noqa: F401  (fails loudly when librt2015.so is missing)
from . import assignments, multi  # noqa: F401
from .host import (  # noqa: F401
    Bounds, Camera, Light, Mesh, Renderer, Vec3, bounds2AABB, loadScene, parseMeshJSON, parseMeshJSON_native, parsePDB, parsePDB_native,
    splitMaterialData, splitMeshData, splitMolData, splitSphereData, splitTriangleData, slabSplitMeshData, slabSplitMolData, toNormalArray,
    toPosArray, write_png,
)

__all__ = ["lib", "multi", "assignments", "Bounds", "Camera", "Light", "Mesh", "Renderer", "Vec3", "bounds2AABB", "loadScene", "parseMeshJSON", "parseMeshJSON_native", "parsePDB_native",
           "parsePDB", "splitMaterialData", "splitMeshData", "splitMolData", "splitSphereData", "splitTriangleData", "slabSplitMeshData", "slabSplitMolData", "toPosArray", "toNormalArray", "write_png"]
