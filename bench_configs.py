#!/usr/bin/env python3
"""bench_configs.py -- the deterministic configs of BASELINE.json (1-4): grid-traversal ms/frame on one
B200 next to the reference kernels on the host cores.  Supplementary to bench.py (which measures the
headline config 5); writes one JSON line per config and, with --out, a JSON file for profiles/.

    python bench_configs.py [--out profiles/rNN_configs.json] [--quick]

GPU time = CUDA events on the context's stream around the assignment's kernel sequence (uploads, grid
build and read-back excluded; grid build reported separately), best of --reps.  CPU time = the same
sequence through oracle/_ref (the reference's code.cl compiled by g++ -O2 -fopenmp), on a smaller frame
where the full one would not fit the time/memory budget -- compared per ray slot.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import synth  # noqa: E402

PKG = "2015-raytracing_b200"


def best_gpu(fn, reps):
    ms = []
    for _ in range(reps + 1):
        ms.append(fn()[-1])
    return min(ms[1:])


def cpu_time(fn, reps=2):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="small frames (smoke run)")
    args = ap.parse_args()
    rt = importlib.import_module(PKG)
    A = rt.assignments
    from oracle import host as OH
    from oracle import refcl as OR
    olib = OR.load_best()
    ctx = rt.lib.Context(0)
    info = ctx.device_info()
    W, H = (480, 270) if args.quick else (1920, 1080)
    cw, ch = (240, 135) if args.quick else (480, 270)     # CPU frame
    results = []

    def emit(name, px_gpu, gpu_ms, px_cpu, cpu_ms, **extra):
        r = {"config": name, "gpu_ms_per_frame": round(gpu_ms, 4), "gpu_ns_per_slot": round(gpu_ms * 1e6 / px_gpu, 4),
             "cpu_ms_per_frame": round(cpu_ms, 3), "cpu_ns_per_slot": round(cpu_ms * 1e6 / px_cpu, 3), "cpu_cores": int(olib.num_threads()),
             "cpu_kind": olib.kind, "speedup_per_slot": round((cpu_ms / px_cpu) / (gpu_ms / px_gpu), 1)}
        r.update(extra)
        results.append(r)
        print(json.dumps(r))
        sys.stdout.flush()

    # ---- config 1: A01 single sphere 512x512
    n = 128 if args.quick else 512
    g = best_gpu(lambda: A.a01_compute(ctx, n, n, timing=True), args.reps)
    c = cpu_time(lambda: OR.a01_render(olib, n, n))
    emit("1: A01 single sphere %dx%d" % (n, n), n * n, g, n * n, c)

    # ---- config 2: molecule, brute force (A02 fused, A03 two kernels); 9018 atoms + the trailing NaN record
    for n_atoms in ((2000,) if args.quick else (9018, 100000)):
        text = synth.synth_pdb(n_atoms=n_atoms, gap_at=n_atoms // 2)
        mol_p, mol_o = rt.parsePDB(text), OH.parsePDB(text)
        for name, gf, cf in (("A02 raytrace", lambda: A.a02_compute(ctx, mol_p, W, H, timing=True), lambda: OR.a02_render(olib, mol_o, cw, ch)),
                             ("A03 initTrace+molTrace", lambda: A.a03_compute(ctx, mol_p, W, H, timing=True), lambda: OR.a03_render(olib, mol_o, cw, ch))):
            if n_atoms > 20000 and name.startswith("A03"):
                continue
            g = best_gpu(gf, args.reps)
            c = cpu_time(cf, 1)
            tests = W * H * mol_p["size"]
            emit("2: %s, %d spheres, %dx%d" % (name, mol_p["size"], W, H), W * H, g, cw * ch, c, sphere_tests_per_s=round(tests / (g * 1e-3), 0),
                 fp32_issue_frac=round(tests * 18 / (g * 1e-3) / (info["sm_count"] * 128 * 1.965e9), 3))

    # ---- A04 / A05 / A06 on the shapes their demos use: a molecule and a ~1 k-triangle mesh
    text = synth.synth_pdb(n_atoms=2000 if args.quick else 9018, gap_at=1000)
    mol_p, mol_o = rt.parsePDB(text), OH.parsePDB(text)
    model = synth.synth_mesh(31, 16, seed=2015)
    md_p, md_o = rt.parseMeshJSON(model), OH.parseMeshJSON(model)
    for name, gf, cf in (
            ("A04 computeBoth (brute force)", lambda: A.a04_compute(ctx, W, H, molData=mol_p, meshData=md_p, timing=True),
             lambda: OR.a04_render(olib, cw, ch, molData=mol_o, meshData=md_o)),
            ("A05 computeBoth (+ boxes)", lambda: A.a05_compute(ctx, W, H, molData=mol_p, meshData=md_p, timing=True),
             lambda: OR.a05_render(olib, cw, ch, molData=mol_o, meshData=md_o)),
            ("A06 computeBoth (x slabs, n 5)", lambda: A.a06_compute(ctx, W, H, 5, molData=mol_p, meshData=md_p, timing=True),
             lambda: OR.a06_render(olib, cw, ch, 5, molData=mol_o, meshData=md_o))):
        g = best_gpu(gf, args.reps)
        c = cpu_time(cf, 1)
        emit("f4: %s, %d spheres + %d triangles, %dx%d" % (name, mol_p["size"], md_p["nTriangles"], W, H), W * H, g, cw * ch, c)

    # ---- config 3: A07 grid primary rays: small mesh and the synthetic 1 M-triangle mesh
    for (mu, mv, ns) in (((31, 16, 10),) if args.quick else ((31, 16, 10), (1000, 500, 128))):
        model = synth.synth_mesh(mu, mv, seed=2015)
        md_p, md_o = rt.parseMeshJSON(model), OH.parseMeshJSON(model)
        t0 = time.perf_counter()
        gg = rt.splitMeshData(ctx, md_p, ns)
        ctx.finish()
        build_ms = (time.perf_counter() - t0) * 1e3
        import ctypes as C
        rt.lib.dll.rt_grid_release(ctx.h, C.byref(gg))
        g = best_gpu(lambda: A.a07_compute(ctx, W, H, ns, meshData=md_p, timing=True), args.reps)
        c = cpu_time(lambda: OR.a07_render(olib, cw, ch, ns, meshData=md_o), 1)
        emit("3: A07 initTrace+meshTrace, %d triangles, n_slabs %d, %dx%d" % (md_p["nTriangles"], ns, W, H), W * H, g, cw * ch, c,
             gpu_grid_build_ms_incl_upload=round(build_ms, 2), mrays_per_s=round(W * H / (g * 1e-3) / 1e6, 1))

    # ---- config 4: A08 (rpp 1) and A09 (rpp 100 default; CPU at rpp 4) on the synthetic Cornell scene, n_slabs 5
    tmp = tempfile.mkdtemp(prefix="rt_cfg_")
    path = synth.write_scene(tmp, n_lights=2, with_sphere=True, with_mesh=False)
    sc8_p, sc8_o = rt.loadScene(path, W, H, assignment=8), OH.loadScene(path, cw, ch, assignment=8)
    g = best_gpu(lambda: A.a08_render(ctx, sc8_p, W, H, 5, timing=True), args.reps)
    c = cpu_time(lambda: OR.a08_render(olib, sc8_o, cw, ch, 5))
    emit("4a: A08 render (2 point lights), %dx%d, rpp 1" % (W, H), W * H, g, cw * ch, c)
    g = best_gpu(lambda: A.a089_render_fused(ctx, sc8_p, W, H, 8, 1, 5, timing=True), args.reps)
    emit("4a: A08 frame in one launch (rt_a089_render_frame), %dx%d, rpp 1" % (W, H), W * H, g, cw * ch, c)
    rpp = 16 if args.quick else 100
    sc9_p, sc9_o = rt.loadScene(path, W, H, assignment=9), OH.loadScene(path, cw, ch, assignment=9)
    g = best_gpu(lambda: A.a09_render(ctx, sc9_p, W, H, rpp, 5, timing=True), max(1, args.reps - 1))
    c = cpu_time(lambda: OR.a09_render(olib, sc9_o, cw, ch, 4, 5), 1)
    emit("4b: A09 render (thin lens), %dx%d, rpp %d (CPU: rpp 4)" % (W, H, rpp), W * H * rpp, g, cw * ch * 4, c)
    g = best_gpu(lambda: A.a089_render_fused(ctx, sc9_p, W, H, 9, rpp, 5, timing=True), max(1, args.reps - 1))
    emit("4b: A09 frame in one launch (rt_a089_render_frame), %dx%d, rpp %d (CPU: rpp 4)" % (W, H, rpp), W * H * rpp, g, cw * ch * 4, c)
    ctx.close()
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"device": info, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
