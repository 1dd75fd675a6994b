"""Randomised-scene parity campaign for the Assignment-10 path: random XML scenes (spheres, axis-aligned and free
triangles incl. degenerate ones, 1-3 disk lights with un-normalised normals, optional <mesh> grids, global n_slabs 1-3,
random camera / lens) rendered for two progressive passes by the wavefront pipeline (a subset also by the megakernel and by the
kernel-by-kernel reference schedule) and compared with the oracle bit
for bit: accumulation image, seed buffer (= RNG draw counts and order), pixels and valid-ray counts.  The fixed
fixtures cover the reference's own scenes; this covers what they do not: open scenes, empty sets, boxes that fail the
shared-far-plane check next to boxes that pass it, multi-cell sphere/triangle sets next to 1-cell ones, rays that leave
the scene, lights inside geometry."""
import os

import numpy as np
import pytest

import synth
from oracle import host as OH
from oracle import refcl as OR

pytestmark = pytest.mark.gpu

COLS, ROWS = 40, 30


def random_scene_xml(rng, with_mesh):
    r = lambda lo, hi: float(np.float32(rng.uniform(lo, hi)))   # noqa: E731  (fp32-representable decimals keep %r short)
    s = "﻿<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<scene>\n<!-- random scene %d -->\n" % int(rng.integers(1 << 30))
    eye = (r(-0.5, 0.5), r(-0.2, 0.6), r(2.0, 3.0))
    s += "<camera>%s%s%s<fov>%r</fov><focal_length>%r</focal_length><lens_diameter>%r</lens_diameter></camera>\n" % (
        synth._v("eye", *eye), synth._v("lookAt", r(-0.2, 0.2), r(-0.3, 0.1), 0.0), synth._v("vup", 0.0, 1.0, 0.0), r(40, 70), r(1.5, 3.0), r(0.0, 0.08))
    for _ in range(int(rng.integers(1, 4))):
        s += synth._light((r(-0.8, 0.8), r(0.2, 0.9), r(-0.8, 0.8)), (r(-1, 1), r(-1.5, -0.2), r(-1, 1)), (r(5, 30), r(5, 30), r(5, 30)), r(0.03, 0.2))
    mats = ["m%d" % k for k in range(4)]
    for m in mats:
        s += synth._mat(m, r(0.1, 0.95), r(0.1, 0.95), r(0.1, 0.95))
    pick = lambda: mats[int(rng.integers(len(mats)))]   # noqa: E731
    for _ in range(int(rng.integers(0, 5))):
        s += "<sphere>%s<radius>%r</radius><matId>%s</matId></sphere>\n" % (synth._v("center", r(-0.8, 0.8), r(-0.8, 0.5), r(-0.8, 0.8)), r(0.05, 0.35), pick())
    shape = int(rng.integers(0, 4))
    if shape >= 1:   # floor (+ walls): axis-aligned quads, the layout the far-plane and 1-cell fast paths are made for
        A = r(0.85, 1.0)
        s += synth._quad((-1, -A, 1.5), (1, -A, 1.5), (1, -A, -1), (-1, -A, -1), (0.0, 1.0, 0.0), pick())
        if shape >= 2:
            s += synth._quad((-1, -1, -A), (1, -1, -A), (1, 1, -A), (-1, 1, -A), (0.0, 0.0, 1.0), pick())
            s += synth._quad((-A, -1, 1.5), (-A, -1, -1), (-A, 1, -1), (-A, 1, 1.5), (1.0, 0.0, 0.0), pick())
        if shape >= 3:
            s += synth._quad((A, -1, -1), (A, -1, 1.5), (A, 1, 1.5), (A, 1, -1), (-1.0, 0.0, 0.0), pick())
            s += synth._quad((-1, A, -1), (1, A, -1), (1, A, 1.5), (-1, A, 1.5), (0.0, -1.0, 0.0), pick())
    for _ in range(int(rng.integers(0, 7))):   # free triangles, some degenerate
        p = [(r(-0.9, 0.9), r(-0.9, 0.9), r(-0.9, 0.9)) for _ in range(3)]
        if rng.random() < 0.15:
            p[2] = p[1]
        n = (r(-1, 1), r(-1, 1), r(0.1, 1))
        s += synth._tri(p[0], p[1], p[2], n, pick())
    if with_mesh:
        s += ("<mesh><file>./tri/synth.json</file><nslabs>%d</nslabs><normalize>%s</normalize>%s%s<matId>%s</matId></mesh>\n"
              % (int(rng.integers(1, 13)), "yes" if rng.random() < 0.7 else "no", synth._v("scale", r(0.3, 0.8), r(0.3, 0.8), r(0.3, 0.8)),
                 synth._v("translate", r(-0.3, 0.3), r(-0.5, 0.2), r(-0.3, 0.3)), pick()))
    return s + "</scene>\n"


CASES = [(c, 0) for c in range(32)] + [(c, 2) for c in range(0, 32, 5)] + [(c, 1) for c in range(1, 32, 7)]


@pytest.mark.parametrize("case,mode", CASES)
def test_random_scene_matches_oracle(rt, oracle_lib, tmp_path, case, mode):
    rng = np.random.Generator(np.random.PCG64(7000 + case))
    with_mesh = case % 2 == 0
    rpp = (4, 9, 1, 16)[case % 4] if case % 4 != 2 else 4   # stratified grids; rpp == 1 has its own (serial-order) test
    n_slabs = (1, 1, 2, 3)[(case // 4) % 4]
    d = tmp_path / "scenes"
    d.mkdir()
    (tmp_path / "tri").mkdir()
    path = str(d / "random.xml")
    with open(path, "w", encoding="utf-8") as f:
        f.write(random_scene_xml(rng, with_mesh))
    synth.mesh_to_json_file(synth.synth_mesh(int(rng.integers(6, 20)), int(rng.integers(4, 12)), seed=case), os.path.join(str(tmp_path), "tri", "synth.json"))
    o_scene, p_scene = OH.loadScene(path, COLS, ROWS), rt.loadScene(path, COLS, ROWS)
    total = COLS * ROWS * rpp
    seeds0 = OR.make_seeds(total, 900 + case)
    prep = OR.prepare_a10(o_scene, n_slabs)
    st = OR.A10State(total, seeds0)
    oracle_lib.a10_initAcu(st.acu, total)
    cam = o_scene["camera"].toFloat32Array()
    r = rt.Renderer(p_scene, COLS, ROWS, rpp, n_slabs=n_slabs, mode=mode)
    r.preRender(seeds0)
    try:
        closest = anyh = 0
        for p in range(2):
            pix_o = OR.a10_execute_render(oracle_lib, st, prep, cam, COLS, ROWS, rpp, o_scene["focal_length"], o_scene["lens_diameter"])
            pix = r.executeRender()
            ref = np.zeros((COLS * ROWS, 4), np.float32)
            for k in range(rpp):
                ref += st.acu.reshape(COLS * ROWS, rpp, 4)[:, k]
            assert np.array_equal(r.seeds(), st.seeds), "case %d pass %d: RNG streams differ" % (case, p)
            assert np.array_equal(r.accum().view(np.uint32), ref.view(np.uint32)), "case %d pass %d: accumulation image differs" % (case, p)
            assert np.array_equal(pix, pix_o.reshape(pix.shape)), "case %d pass %d: pixels differ" % (case, p)
            s = r.stats()
            closest += s["closest_rays"]
            anyh += s["any_rays"]
            assert (closest, anyh) == (st.n_closest, st.n_any), "case %d pass %d: valid-ray counts differ" % (case, p)
    finally:
        r.postRender()


@pytest.mark.parametrize("case", range(12))
def test_random_scene_a08_a09_match_oracle(rt, gpu_ctx, ref_lib, tmp_path, case):
    """The same random scenes through the deterministic frames of Assignments 8 and 9 (point lights, shadow rays, thin
    lens; launcher by launcher like their render()), n_slabs 1 / 3 / 5: hit ids, hit distances, float image and pixels."""
    rng = np.random.Generator(np.random.PCG64(8000 + case))
    d = tmp_path / "scenes"
    d.mkdir()
    path = str(d / "random.xml")
    with open(path, "w", encoding="utf-8") as f:
        f.write(random_scene_xml(rng, False))
    n_slabs = (1, 3, 5)[case % 3]
    for a in (8, 9):
        o_scene, p_scene = OH.loadScene(path, COLS, ROWS, assignment=a), rt.loadScene(path, COLS, ROWS, assignment=a)
        if not (len(o_scene["spheres"]) or len(o_scene["triangles"])):
            continue
        if a == 8:
            acu_o, pix_o, st = OR.a08_render(ref_lib, o_scene, COLS, ROWS, n_slabs)
            acu, pix, matid, maxt = rt.assignments.a08_render(gpu_ctx, p_scene, COLS, ROWS, n_slabs)
        else:
            acu_o, pix_o, st = OR.a09_render(ref_lib, o_scene, COLS, ROWS, 4, n_slabs)
            acu, pix, matid, maxt = rt.assignments.a09_render(gpu_ctx, p_scene, COLS, ROWS, 4, n_slabs)
        want_id = st["pois"]["matId"].astype(np.int32)
        assert np.array_equal(matid, want_id), "A%02d case %d: hit ids" % (a, case)
        hit = want_id >= 0
        assert np.array_equal(maxt[hit].view(np.uint32), st["rays"]["maxt"][hit].view(np.uint32)), "A%02d case %d: hit distances" % (a, case)
        assert np.array_equal(acu.view(np.uint32), np.ascontiguousarray(acu_o, dtype=np.float32).view(np.uint32)), "A%02d case %d: float image" % (a, case)
        assert np.array_equal(pix.reshape(-1, 4), np.asarray(pix_o).reshape(-1, 4)), "A%02d case %d: pixels" % (a, case)
        acu_f, pix_f, matid_f, maxt_f = rt.assignments.a089_render_fused(gpu_ctx, p_scene, COLS, ROWS, a, 4, n_slabs)
        assert np.array_equal(acu_f.view(np.uint32), acu.view(np.uint32)) and np.array_equal(matid_f, matid) and np.array_equal(pix_f, pix), \
            "A%02d case %d: one-launch frame differs from the launcher sequence" % (a, case)
        assert np.array_equal(maxt_f[hit].view(np.uint32), maxt[hit].view(np.uint32))
