"""GPU parity tests of Assignments 4-6 through the C ABI: brute force over spheres and a
triangle soup (A04), the same behind bounding boxes (A05), 1-D slabs along x built on the GPU (A06) -- against the
oracle (the reference's own code.cl text, oracle/_ref) on seeded synthetic inputs, plus size-independent
properties at the full 1920x1080 frame."""
import ctypes as C

import numpy as np
import pytest

import synth
from oracle import host as OH
from oracle import refcl as OR

pytestmark = pytest.mark.gpu

COLS, ROWS = 320, 240


@pytest.fixture(scope="module")
def inputs(rt):
    text = synth.synth_pdb(n_atoms=300, gap_at=120, seed=7)      # size = records + 1: one all-NaN sphere (quirk Q13)
    model = synth.synth_mesh(24, 12, seed=11)
    return {"mol_p": rt.parsePDB(text), "mol_o": OH.parsePDB(text), "mesh_p": rt.parseMeshJSON(model), "mesh_o": OH.parseMeshJSON(model)}


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


FLOWS = [("compute", True, False), ("computeTri", False, True), ("computeBoth", True, True)]


@pytest.mark.parametrize("flow,use_mol,use_mesh", FLOWS)
def test_a04_a05_match_oracle(rt, gpu_ctx, ref_lib, inputs, flow, use_mol, use_mesh):
    kw_p = {"molData": inputs["mol_p"] if use_mol else None, "meshData": inputs["mesh_p"] if use_mesh else None}
    kw_o = {"molData": inputs["mol_o"] if use_mol else None, "meshData": inputs["mesh_o"] if use_mesh else None}
    for gpu, cpu in ((rt.assignments.a04_compute, OR.a04_render), (rt.assignments.a05_compute, OR.a05_render)):
        pix, maxt = gpu(gpu_ctx, COLS, ROWS, **kw_p)
        opix, orays = cpu(ref_lib, COLS, ROWS, **kw_o)
        assert np.array_equal(_bits(maxt), _bits(orays["maxt"])), "%s %s: hit distances / hit set" % (gpu.__name__, flow)
        assert np.array_equal(pix, opix), "%s %s: pixels" % (gpu.__name__, flow)
        assert np.isfinite(maxt).sum() > 1000   # the frame is not empty


@pytest.mark.parametrize("n_slabs", [1, 3, 5, 16])
@pytest.mark.parametrize("flow,use_mol,use_mesh", FLOWS)
def test_a06_matches_oracle(rt, gpu_ctx, ref_lib, inputs, flow, use_mol, use_mesh, n_slabs):
    kw_p = {"molData": inputs["mol_p"] if use_mol else None, "meshData": inputs["mesh_p"] if use_mesh else None}
    kw_o = {"molData": inputs["mol_o"] if use_mol else None, "meshData": inputs["mesh_o"] if use_mesh else None}
    pix, maxt = rt.assignments.a06_compute(gpu_ctx, COLS, ROWS, n_slabs, **kw_p)
    opix, orays, prep = OR.a06_render(ref_lib, COLS, ROWS, n_slabs, **kw_o)
    # the slab lists themselves (GPU build against the JS-order restatement)
    if use_mol:
        g = rt.slabSplitMolData(gpu_ctx, inputs["mol_p"], n_slabs)
        d = rt.host.DeviceGrid(gpu_ctx, g, np.zeros(8, np.float32), cells=n_slabs)
        m = prep["mol"]
        try:
            assert np.array_equal(d.box_size(), m["box"]) and np.array_equal(_bits(d.prim()), _bits(m["atoms"]))
            assert np.array_equal(d.matid(), m["index"])
        finally:
            rt.lib.dll.rt_grid_release(gpu_ctx.h, C.byref(g))
    if use_mesh:
        g = rt.slabSplitMeshData(gpu_ctx, inputs["mesh_p"], n_slabs)
        d = rt.host.DeviceGrid(gpu_ctx, g, np.zeros(8, np.float32), cells=n_slabs)
        t = prep["mesh"]
        try:
            assert np.array_equal(d.box_size(), t["box"]) and np.array_equal(_bits(d.prim()), _bits(t["pos"]))
            assert np.array_equal(_bits(d.normal()), _bits(t["normal"])) and np.array_equal(d.matid(), t["index"])
        finally:
            rt.lib.dll.rt_grid_release(gpu_ctx.h, C.byref(g))
    assert np.array_equal(_bits(maxt), _bits(orays["maxt"])), "A06 %s n=%d: hit distances / hit set" % (flow, n_slabs)
    assert np.array_equal(pix, opix), "A06 %s n=%d: pixels" % (flow, n_slabs)


def test_full_frame_properties(rt, gpu_ctx, inputs):
    """1920x1080, no oracle: (1) with all geometry strictly inside its boxes the bounding boxes of A05 only cull, so
    A05's image and hit distances equal A04's; (2) the fused A04 `raytrace` kernel shows the same spheres as
    initTrace + molTrace wherever the unclamped shade is in range; (3) A06 finds the same nearest primitive as A05
    (its sphere arithmetic differs by the explicit mad, so distances are compared to 1e-4 relative)."""
    A = rt.assignments
    W, H = 1920, 1080
    mol, mesh = inputs["mol_p"], inputs["mesh_p"]
    p4, t4 = A.a04_compute(gpu_ctx, W, H, molData=mol, meshData=mesh)
    p5, t5 = A.a05_compute(gpu_ctx, W, H, molData=mol, meshData=mesh)
    hit4 = np.isfinite(t4)      # A04 rays start as (0, inf): finite maxt <=> a primitive was hit
    assert hit4.sum() > 100000
    # A05 clips the primary ray to the scene box and tests each set's box first; a ray that grazes a box face exactly
    # where a primitive touches it may be culled by float rounding, hence a (tiny) budget instead of equality
    differs = np.any(p4.reshape(-1, 4) != p5.reshape(-1, 4), axis=1) | (_bits(t4) != _bits(t5))
    assert (differs & hit4).sum() <= 1e-5 * hit4.sum(), "A05 differs from A04 on %d of %d hit pixels" % ((differs & hit4).sum(), hit4.sum())
    assert not np.any(p5.reshape(-1, 4)[~hit4][:, :3]), "A05 lit a pixel A04 left black"
    pm, tm = A.a04_compute(gpu_ctx, W, H, molData=mol)
    fused = A.a04_raytrace(gpu_ctx, mol, W, H)
    lit = pm.reshape(-1, 4)[:, :3].sum(axis=1) > 0          # clamped shade > 0 <=> the unclamped one is the same number
    assert np.array_equal(fused.reshape(-1, 4)[lit], pm.reshape(-1, 4)[lit])
    # both start from the same initTrace (finite maxt = exit of the scene box) and lower maxt to the nearest primitive
    p6, t6 = A.a06_compute(gpu_ctx, W, H, 5, molData=mol, meshData=mesh)
    box = np.isfinite(t5)
    assert np.array_equal(box, np.isfinite(t6))
    off = np.abs(t5[box] - t6[box]) > 1e-4 * np.abs(t5[box]).max()
    assert off.mean() < 1e-4, "A06 and A05 disagree on the nearest primitive of %d pixels" % off.sum()
