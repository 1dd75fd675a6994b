#!/usr/bin/env python3
"""bench_configs.py -- the deterministic configs of BASELINE.json (1-4) on the inputs BASELINE.md section 3 names:
grid-traversal ms/frame on one B200 with the roofline that bounds each, next to the reference kernels on the host cores.
Supplementary to bench.py (which measures the headline config 5 and attaches the compact form of this list to its JSON
line); writes one JSON line per config and, with --out, a JSON file for profiles/.

    python bench_configs.py [--out profiles/rNN_configs.json] [--quick] [--no-cpu]

Inputs (all from tests/golden/, the reference's own files in neutral form -- /root/reference does not exist on the GPU box):
  1  A01 512x512                                    2  mol/3IZ4.pdb (9018 atoms + the trailing NaN record), synthetic 1e4 / 1e5
  3  tri/{teapot, house, house_of_parliament}.json at n_slabs 2 / 10 / 32; synthetic 1 M triangles at 64 / 128 / 256
  4  every A08 scene (rpp 1) and every A09 scene (rpp 100, the default of A09/code.js:232-235), n_slabs 5
GPU time = CUDA events on the context's stream around the assignment's kernel sequence (uploads, grid build and read-back
excluded; grid build reported separately), best of --reps.  Roofline: config 2 against the FP32 issue rate (18 operations
per sphere test); configs 3 and 4 the ALGORITHMIC bytes of BASELINE.md section 3 -- 48 per alive ray + 8 C + B T + H (...)
with C / T / H counted on the device by an instrumented run of the same launchers (rt_set_walk_totals; equal to the
oracle's counters, tests/test_gpu_gates.py) plus the streaming kernels' bytes -- over the kernel time, against the
measured HBM peak.  CPU time = the same sequence through oracle/_ref (the reference's code.cl compiled by g++ -O2 -fopenmp)
on a smaller frame where the full one would not fit the time / memory budget -- compared per ray slot."""
from __future__ import annotations

import argparse
import ctypes as C
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

import golden_io as G  # noqa: E402
import synth  # noqa: E402

PKG = "2015-raytracing_b200"


def best_gpu(fn, reps):
    ms = []
    for _ in range(reps + 1):
        ms.append(fn()[-1])
    return min(ms[1:])


def cpu_time(fn, reps=1):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


class Totals:
    """rt_set_walk_totals: device counters the grid-walk launchers add to while switched on."""
    NAMES = ("alive", "walks", "cells", "tests", "hits", "front")

    def __init__(self, rt, ctx):
        self.rt, self.ctx = rt, ctx
        self.d = ctx.alloc(64)

    def measure(self, fn):
        self.ctx.call("rt_buffer_fill", self.d, 0, 64)
        self.ctx.call("rt_set_walk_totals", self.d)
        try:
            fn()
        finally:
            self.ctx.call("rt_set_walk_totals", None)
        v = self.ctx.download(self.d, np.uint64, 8)
        return {k: int(v[i]) for i, k in enumerate(self.NAMES)}

    def close(self):
        self.ctx.free(self.d)


def run_configs(rt, ctx, compact=False, with_cpu=True, reps=3, quick=False, big_mesh=None, emit=None, hbm_peak=6548.8):
    """Runs the config list and returns the result dicts.  compact = the subset bench.py attaches to its line (one input per
    config, no CPU leg, ~15 s)."""
    A = rt.assignments
    olib = OH = OR = None
    if with_cpu:
        from oracle import host as OH
        from oracle import refcl as OR
        olib = OR.load_best()
        olib.set_num_threads(os.cpu_count() or 1)
    info = ctx.device_info()
    W, H = (480, 270) if quick else (1920, 1080)
    cw, ch = (240, 135) if quick else (480, 270)     # CPU frame
    fp32_peak = info["sm_count"] * 128 * 1.965e9
    tot = Totals(rt, ctx)
    results = []
    tmp = tempfile.mkdtemp(prefix="rt_cfg_")

    def out(name, slots_gpu, gpu_ms, slots_cpu=None, cpu_ms=None, **extra):
        r = {"config": name, "gpu_ms_per_frame": round(gpu_ms, 4), "gpu_ns_per_slot": round(gpu_ms * 1e6 / slots_gpu, 4)}
        if cpu_ms is not None:
            r.update({"cpu_ms_per_frame": round(cpu_ms, 3), "cpu_ns_per_slot": round(cpu_ms * 1e6 / slots_cpu, 3), "cpu_cores": int(olib.num_threads()),
                      "cpu_kind": olib.kind, "speedup_per_slot": round((cpu_ms / slots_cpu) / (gpu_ms / slots_gpu), 1)})
        r.update(extra)
        results.append(r)
        if emit:
            emit(r)

    def hbm(bytes_, ms):
        gbs = bytes_ / (ms * 1e-3) / 1e9
        frac = gbs / hbm_peak
        # an ALGORITHMIC rate (reference layout: 48 B per triangle test ...) above the HBM peak means the scene is cache resident
        # and the kernel is bound by instruction issue, not by memory
        bound = "hbm (algorithmic bytes, BASELINE.md 3)" if frac <= 1.0 else "instruction issue (cache-resident scene: the algorithmic rate exceeds the HBM peak)"
        return {"algorithmic_bytes": int(bytes_), "achieved_gbs": round(gbs, 1), "hbm_frac": round(frac, 4), "bound": bound}

    # ---- config 1: A01 single sphere 512x512 (latency-bound; no roofline claim)
    n = 128 if quick else 512
    g = best_gpu(lambda: A.a01_compute(ctx, n, n, timing=True), reps)
    c = cpu_time(lambda: OR.a01_render(olib, n, n), 2) if with_cpu else None
    out("1: A01 single sphere %dx%d" % (n, n), n * n, g, n * n, c, bound="launch latency")

    # ---- config 2: molecule, brute force (A02 fused, A03 two kernels): 3IZ4 = 9018 atoms + the trailing NaN record
    fx = G.load("mol_3IZ4")
    mols = [("mol/3IZ4.pdb", G.pdb_text(fx["serial"], fx["elem"], fx["xyz"]))]
    if not compact and not quick:
        mols += [("synthetic 1e4", synth.synth_pdb(n_atoms=10000, gap_at=5000)), ("synthetic 1e5", synth.synth_pdb(n_atoms=100000, gap_at=50000))]
    for label, text in mols:
        mol_p = rt.parsePDB(text)
        mol_o = OH.parsePDB(text) if with_cpu else None
        forms = [("A02 raytrace", lambda: A.a02_compute(ctx, mol_p, W, H, timing=True), (lambda: OR.a02_render(olib, mol_o, cw, ch)))]
        if mol_p["size"] <= 20000 and not compact:
            forms.append(("A03 initTrace+molTrace", lambda: A.a03_compute(ctx, mol_p, W, H, timing=True), (lambda: OR.a03_render(olib, mol_o, cw, ch))))
        for name, gf, cf in forms:
            g = best_gpu(gf, reps)
            c = cpu_time(cf) if with_cpu else None
            tests = W * H * mol_p["size"]
            out("2: %s, %s (%d spheres), %dx%d" % (name, label, mol_p["size"], W, H), W * H, g, cw * ch, c,
                sphere_tests_per_s=round(tests / (g * 1e-3), 0), fp32_issue_frac=round(tests * 18 / (g * 1e-3) / fp32_peak, 3), bound="FP32 issue")

    # ---- config 3: A07 grid primary rays on the reference's meshes and the synthetic 1 M-triangle mesh
    def a07(label, md_p, md_o, ns):
        t0 = time.perf_counter()
        gg = rt.splitMeshData(ctx, md_p, ns)
        ctx.finish()
        build_ms = (time.perf_counter() - t0) * 1e3
        refs = int(gg.n_refs)
        rt.lib.dll.rt_grid_release(ctx.h, C.byref(gg))
        g = best_gpu(lambda: A.a07_compute(ctx, W, H, ns, meshData=md_p, timing=True), reps)
        t = tot.measure(lambda: A.a07_compute(ctx, W, H, ns, meshData=md_p))
        # initTrace: 48 W + 4 W per pixel; meshTrace: 48 per alive ray + 8 C + 48 T + H (48 normals + 4 maxt + 4 pixel)
        byts = 52 * W * H + 48 * t["alive"] + 8 * t["cells"] + 48 * t["tests"] + 56 * t["hits"]
        c = cpu_time(lambda: OR.a07_render(olib, cw, ch, ns, meshData=md_o)) if with_cpu else None
        out("3: A07 initTrace+meshTrace, %s (%d triangles, %d refs), n_slabs %d, %dx%d" % (label, md_p["nTriangles"], refs, ns, W, H), W * H, g, cw * ch, c,
            gpu_grid_build_ms_incl_upload=round(build_ms, 2), mrays_per_s=round(t["alive"] / (g * 1e-3) / 1e6, 1),
            cells_per_ray=round(t["cells"] / max(t["alive"], 1), 2), tests_per_ray=round(t["tests"] / max(t["alive"], 1), 2),
            hit_frac=round(t["hits"] / max(t["alive"], 1), 4), **hbm(byts, g))

    named = [("tri_teapot", "tri/teapot.json"), ("tri_house", "tri/house.json"), ("tri_house_of_parliament", "tri/house_of_parliament.json")]
    for fxname, label in (named[2:] if compact else named):
        fx = G.load(fxname)
        m = G.meshes_of(fx)[0]
        path = os.path.join(tmp, fxname + ".json")
        with open(path, "w") as f:
            f.write(G.mesh_json_text(m["positions"], m["normals"], m["materialIndices"], m["materials"]))
        md_p = rt.parseMeshJSON(path)
        md_o = OH.parseMeshJSON(path) if with_cpu else None
        for ns in ((32,) if compact else ((10,) if quick else (2, 10, 32))):
            a07(label, md_p, md_o, ns)
    if not quick:
        model = None
        if big_mesh is None:
            model = synth.synth_mesh(1000, 500, seed=2015)
            big_mesh = rt.parseMeshJSON(model)
        md_o = None
        if with_cpu:
            md_o = OH.parseMeshJSON(model if model is not None else synth.synth_mesh(1000, 500, seed=2015))
        for ns in ((128,) if compact else (64, 128, 256)):
            a07("synthetic mesh", big_mesh, md_o, ns)

    # ---- config 4: A08 (rpp 1) and A09 (rpp 100, the default) on the reference's scenes, n_slabs 5
    def a089(fxname, assignment, rpp):
        fx = G.load(fxname)
        d = os.path.join(tmp, fxname)
        os.makedirs(d, exist_ok=True)
        path = G.materialize_scene(fx["tree"], [], d)
        sc = rt.loadScene(path, W, H, assignment=assignment)
        nl = len(sc["lights"])
        slots = W * H * rpp
        if assignment == 8:
            seq = lambda timing=False: A.a08_render(ctx, sc, W, H, 5, timing=timing)   # noqa: E731
        else:
            seq = lambda timing=False: A.a09_render(ctx, sc, W, H, rpp, 5, timing=timing)   # noqa: E731
        g_seq = best_gpu(lambda: seq(True), max(1, reps - 1))
        g_one = best_gpu(lambda: A.a089_render_fused(ctx, sc, W, H, assignment, rpp, 5, timing=True), max(1, reps - 1))
        t = tot.measure(seq)
        # per slot: initTrace 48 + 16 + 4 W; per light initShadowTrace 48 R + 48 W and sceneRender 48 + 48 + 16 R + 32 RMW; copyToPixel
        # 16 rpp R + 4 W per pixel; traversal 48 per alive (ray, set) query + 8 C + B T (B = 48: triangles dominate these scenes; spheres
        # are 16) + per closest hit 48 normals + 48 Poi + 4 maxt, per shadow walk 8
        byts = slots * (68 + nl * (96 + 144)) + W * H * (16 * rpp + 4) + 48 * t["alive"] + 8 * t["cells"] + 48 * t["tests"] + 100 * t["hits"]
        c = None
        cpu_slots = cw * ch * (1 if assignment == 8 else 4)
        if with_cpu:
            sco = OH.loadScene(path, cw, ch, assignment=assignment)
            c = cpu_time((lambda: OR.a08_render(olib, sco, cw, ch, 5)) if assignment == 8 else (lambda: OR.a09_render(olib, sco, cw, ch, 4, 5)))
        name = "4%s: A%02d scenes/%s.xml, %d lights, %dx%d, rpp %d%s" % ("a" if assignment == 8 else "b", assignment, fxname[4:], nl, W, H, rpp,
                                                                         " (CPU: rpp 4)" if (with_cpu and assignment == 9) else "")
        out(name, slots, g_seq, cpu_slots, c, gpu_ms_one_launch_frame=round(g_one, 4), queries=t["alive"],
            cells_per_query=round(t["cells"] / max(t["alive"], 1), 2), tests_per_query=round(t["tests"] / max(t["alive"], 1), 2), **hbm(byts, g_seq))

    a08 = ["a08_cornell"] if compact else G.names("a08_")
    a09 = ["a09_cornell"] if compact else G.names("a09_")
    for fxname in a08:
        a089(fxname, 8, 1)
    for fxname in a09:
        a089(fxname, 9, 16 if quick else 100)
    tot.close()
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="small frames (smoke run)")
    ap.add_argument("--compact", action="store_true", help="the subset bench.py attaches to its JSON line")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rt = importlib.import_module(PKG)
    ctx = rt.lib.Context(0)
    info = ctx.device_info()
    peak = 6548.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def emit(r):
        print(json.dumps(r))
        sys.stdout.flush()

    results = run_configs(rt, ctx, compact=args.compact, with_cpu=not args.no_cpu, reps=args.reps, quick=args.quick, emit=emit, hbm_peak=peak)
    ctx.close()
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"device": info, "hbm_peak_gbs": peak, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
