#!/usr/bin/env python3
"""bench.py -- headline benchmark: path-traced Mrays/s at 1920x1080 (BASELINE.json config 5:
Assign10 path tracing, 256 slots per pixel as a 16x16 stratified lens grid, depth 5, synthetic
1M-triangle <mesh> at nslabs 128 inside a Cornell-style box with two disk lights), split by
samples-per-pixel across the GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one executeRender pass (Assignment10/code.js:1806-1854) over the whole frame.
Prints ONE JSON line (see README/DESIGN.md for the field contract).  The oracle is imported
only for the cpu_baseline / reference legs -- never on the measured product path.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

PKG = "2015-raytracing_b200"
METRIC = "path-traced Mrays/s at 1080p"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cols", type=int, default=1920)
    ap.add_argument("--rows", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=256, help="slots per pixel (perfect square)")
    ap.add_argument("--depth", type=int, default=5)
    ap.add_argument("--mesh-u", type=int, default=1000, help="quads around (triangles = 2*u*v)")
    ap.add_argument("--mesh-v", type=int, default=500)
    ap.add_argument("--nslabs", type=int, default=128)
    ap.add_argument("--lights", type=int, default=2)
    ap.add_argument("--mode", type=int, default=0, help="0 fused (default), 1 reference kernel schedule")
    ap.add_argument("--tile-slots", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact configs 1-4 block (bench_configs.py) appended at N = 1")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the frame the CPU baseline renders (0 = auto)")
    ap.add_argument("--seed", type=int, default=2015)
    ap.add_argument("--profile-passes", type=int, default=0, help="profiling aid (ncu): run this many passes of the workload and exit, no timing legs")
    return ap.parse_args()


# ------------------------------------------------------------------------------ workload
def build_scene(rt, args, tmp):
    import synth
    mesh_json = synth.synth_mesh(args.mesh_u, args.mesh_v, seed=args.seed)
    jmesh_holder = {}

    def loader(_file):
        if "m" not in jmesh_holder:
            jmesh_holder["m"] = rt.parseMeshJSON(mesh_json)
        return jmesh_holder["m"]

    path = synth.write_scene(tmp, n_lights=args.lights, with_sphere=True, with_mesh=True, mesh_nslabs=args.nslabs)
    scene = rt.loadScene(path, args.cols, args.rows, mesh_loader=loader)
    return scene, jmesh_holder.get("m")


def workload_name(args):
    return ("A10 path tracing %dx%d, %d slots/px (stratified %dx%d), depth %d, synthetic %d-triangle mesh nslabs %d, "
            "Cornell-style box + sphere, %d disk lights" % (args.cols, args.rows, args.spp, int(args.spp ** 0.5), int(args.spp ** 0.5),
                                                            args.depth, 2 * args.mesh_u * args.mesh_v, args.nslabs, args.lights))


def config_of(args):
    """`config` of the JSON line -- the same dict in both arms (ours and --impl reference)."""
    slots = args.cols * args.rows * args.spp
    return {"workload": workload_name(args), "cols": args.cols, "rows": args.rows, "slots_per_pixel": args.spp, "depth": args.depth,
            "mesh_triangles": 2 * args.mesh_u * args.mesh_v, "mesh_nslabs": args.nslabs, "lights": args.lights, "seed": args.seed,
            "l2_policy": "inputs larger than L2 (126 MB), no flush needed: every step streams the per-slot seed + accumulation state "
                         "(20 B/slot) and the per-tile ray/hit state (~116 B/slot) of cols*rows*slots_per_pixel/N slots per GPU -- "
                         "%.1f + %.1f GB at N = 1, %.1f + %.1f GB at N = 8" % (slots * 20 / 1e9, slots * 116 / 1e9, slots * 20 / 8e9, slots * 116 / 8e9)}


def make_seeds(total, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    return g.integers(1, 2 ** 31, size=total, dtype=np.int32)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append((float(f[0]), float(f[1])))
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(s[1] for s in self.samples), "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------ CPU legs (oracle)
def oracle_scene(args, tmp):
    """The same synthetic scene through the ORACLE's host restatement of code.js (loadScene,
    parseMeshJSON, split*Data, Mesh transforms) -- no product code and no GPU on this path."""
    import synth
    from oracle import host as OH
    from oracle import refcl as OR
    mesh_json = synth.synth_mesh(args.mesh_u, args.mesh_v, seed=args.seed)
    path = synth.write_scene(tmp, n_lights=args.lights, with_sphere=True, with_mesh=True, mesh_nslabs=args.nslabs)
    scene = OH.loadScene(path, args.cols, args.rows, mesh_loader=lambda _f: OH.parseMeshJSON(mesh_json))
    return scene, OR.prepare_a10(scene, 1)


def sample_rows(args, n):
    """`n` pixel rows spread evenly over the WHOLE frame (row i of n sits in the middle of the i-th of n equal bands): the
    CPU arms render a sample whose mix of mesh / wall / background pixels is the frame's, not the expensive centre's."""
    n = max(1, min(args.rows, n))
    return [min(args.rows - 1, int((i + 0.5) * args.rows / n)) for i in range(n)]


def cpu_pass_rows(olib, prep, cam16, args, focal, lens_diam, rows, seed):
    """One executeRender pass of the oracle over each of the pixel rows in `rows` (every kernel is per-slot independent,
    so a row tile is exact).  Returns (rays, seconds) summed over the rows; only the kernel schedule is timed."""
    from oracle import refcl as OR
    cols, rpp = args.cols, args.spp
    total = cols * rpp
    rays, secs = 0, 0.0
    for i, row in enumerate(rows):
        st = OR.A10State(total, make_seeds(total, seed + 1000 * i))
        olib.a10_initAcu(st.acu, total)
        t0 = time.perf_counter()
        OR.a10_execute_render(olib, st, prep, cam16, cols, args.rows, rpp, focal, lens_diam, depth=args.depth, row0=row, nrows=1)
        secs += time.perf_counter() - t0
        rays += st.n_closest + st.n_any
    return rays, secs


# ------------------------------------------------------------------------------ roofline
WALK_KERNEL = {"walk_triangle_closest": "k_walk_pairs<triangle, closest hit>", "walk_triangle_any": "k_walk_pairs<triangle, any hit>",
               "walk_sphere_closest": "k_walk_pairs<sphere, closest hit>", "walk_sphere_any": "k_walk_pairs<sphere, any hit>"}
NCU_WORKLOAD = (1920, 1080, 256, 5, 1000, 500, 128, 2)


def load_ncu_classes(args, world):
    """The committed ncu capture of ONE pass of this workload (profiles/*_ncu_classes.json, written by
    profiles/ncu_classes.py): per kernel class the DRAM bytes and the warp / thread instructions of its launches.  Returned
    only when it belongs to THIS build -- the file carries a hash of the kernel sources -- and to the default workload at
    N = 1 (per-launch figures do not transfer to other tile sizes); otherwise (None, reason)."""
    if (args.cols, args.rows, args.spp, args.depth, args.mesh_u, args.mesh_v, args.nslabs, args.lights) != NCU_WORKLOAD or args.mode != 0:
        return None, "no ncu capture for a non-default workload"
    if world != 1:
        return None, "the ncu capture is per launch at N = 1 (one GPU, never a multi-rank command)"
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    try:
        import ncu_classes
        sha = ncu_classes.source_sha(ROOT)
    except Exception as e:   # sources missing: cannot prove freshness
        return None, "cannot hash the kernel sources (%s)" % e
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for fn in sorted(os.listdir(pdir)):
        if fn.endswith("_ncu_classes.json"):
            try:
                d = json.load(open(os.path.join(pdir, fn)))
            except Exception:
                continue
            if d.get("source_sha") == sha:
                best = (fn, d)
    if best is None:
        return None, "stale: no profiles/*_ncu_classes.json was captured from the current kernel sources (sha %s)" % sha
    return best, None


def roofline_of(r, scene, tclass, kern_ms, args, peak, measured_peak, world, clocks, info):
    """Roofline of the DOMINANT kernel.  Three ceilings, all per launch of that kernel class:
      frac       ALGORITHMIC bytes (the reference layout's bytes for the work done, counted by the instrumented
                 pass on the same inputs, per geometry set) / launch time (CUDA events inside the timed region) / HBM peak;
      frac_dram  MEASURED DRAM bytes (ncu dram__bytes_read + write of the same launches, from the committed capture of this
                 build) / the same launch time / HBM peak;
      issue      warp instructions issued / (SMs x 4 schedulers x clock x time), the lanes those instructions kept busy, and
                 the share of the FP32 lanes doing the algorithm's own arithmetic (5 flops per face cull + 40 per full
                 triangle test, 18 per sphere test).
    `bound` names the ceiling the kernel sits closest to.  The whole step is reported next to it."""
    alg = r.algorithmic_bytes()
    walk = {k: v for k, v in tclass.items() if k.startswith("walk_") and v["launches"]}
    step_ms = kern_ms / max(args.steps, 1)
    step = {"algorithmic_bytes": alg["bytes_per_pass"], "ms": round(step_ms, 3),
            "achieved_gbs": round(alg["bytes_per_pass"] / (step_ms * 1e-3) / 1e9, 2),
            "class_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in tclass.items() if v["launches"]},
            "class_launches_per_step": {k: v["launches"] // args.steps for k, v in tclass.items() if v["launches"]}}
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if measured_peak else "fallback"
    if not walk:   # megakernel / reference schedule: the step is the kernel
        achieved = step["achieved_gbs"]
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                "kernel": "k_pathMega" if args.mode == 2 else "whole pass", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg["bytes_per_pass"], "launch_ms": round(step_ms, 3), "step": step}
    # heavy sets (the ones the queue walkers serve), by primitive kind; rows of the work profile are in set order
    kinds = []
    if len(scene["spheres"]) > 0:
        kinds.append(("sphere", r._grids[len(kinds)]))
    if len(scene["triangles"]) > 0:
        kinds.append(("triangle", r._grids[len(kinds)]))
    for m in scene["meshes"]:
        kinds.append(("triangle", m.grid))
    byts = {k: 0 for k in WALK_KERNEL}
    flops = {k: 0 for k in WALK_KERNEL}
    prof = alg["profile_sets"]
    for row, p, (kind, g) in zip(alg["per_set"], prof, kinds):
        if not g.n_slabs > 1:   # 1-cell sets are intersected inline by the stage kernels
            continue
        # the walker sees only rays that hit the set's AABB: take the ray loads of the others out
        byts["walk_%s_closest" % kind] += row["closest_bytes"] - 48 * (row["closest_queries"] - row["closest_walks"])
        byts["walk_%s_any" % kind] += row["any_bytes"] - 48 * (row["any_queries"] - row["any_walks"])
        if kind == "triangle":   # 5 flops per cull, 40 more for the tests that pass it; the front count [15] covers both ray kinds
            tc, ta = p[4], p[12]
            front = p[15]
            flops["walk_triangle_closest"] += 5 * tc + 40 * front * tc / max(tc + ta, 1)
            flops["walk_triangle_any"] += 5 * ta + 40 * front * ta / max(tc + ta, 1)
        else:
            flops["walk_sphere_closest"] += 18 * p[3]
            flops["walk_sphere_any"] += 18 * p[11]
    ncu, ncu_why = load_ncu_classes(args, world)
    clock_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
    sms = info["sm_count"]
    dom = max(walk, key=lambda k: walk[k]["ms"])
    per = {}
    for k, v in walk.items():
        n_l = v["launches"]
        b_launch = byts[k] * args.steps / n_l
        ms_launch = v["ms"] / n_l
        t = ms_launch * 1e-3
        e = {"launches_per_step": n_l // args.steps, "launch_ms": round(ms_launch, 4), "algorithmic_bytes_per_launch": int(b_launch),
             "achieved_gbs": round(b_launch / t / 1e9, 2), "frac": round(b_launch / t / 1e9 / peak, 4), "share_of_step": round(v["ms"] / kern_ms, 4),
             "fp32_useful_frac": round(flops[k] * args.steps / n_l / (sms * 128 * clock_hz * t), 4)}
        c = ncu[1]["classes"].get(k) if ncu else None
        if c and c["launches"] == e["launches_per_step"]:
            dram = (c["dram_read"] + c["dram_write"]) / c["launches"]
            e.update({"traffic": int(dram), "frac_dram": round(dram / t / 1e9 / peak, 4),
                      "warp_inst_per_launch": int(c["warp_inst"] / c["launches"]),
                      "issue_frac": round(c["warp_inst"] / c["launches"] / (sms * 4 * clock_hz * t), 4),
                      "lanes_per_inst": round(c["thread_inst"] / max(c["warp_inst"], 1), 2),
                      "lane_frac": round(c["thread_inst"] / c["launches"] / (sms * 4 * 32 * clock_hz * t), 4)})
        per[k] = e
    d = per[dom]
    # which ceiling is the kernel closest to?  (issue slots vs measured DRAM traffic; the algorithmic rate is a contract figure)
    bound = "hbm"
    if "issue_frac" in d and d["issue_frac"] > d.get("frac_dram", 0.0):
        bound = "issue"
    out = {"bound": bound, "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": d["frac"],
           "traffic": d.get("traffic"), "frac_dram": d.get("frac_dram"),
           "kernel": WALK_KERNEL[dom], "peak_source": peak_src,
           "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"], "launch_ms": d["launch_ms"],
           "share_of_step": d["share_of_step"],
           "issue": {k: d.get(k) for k in ("issue_frac", "lanes_per_inst", "lane_frac", "fp32_useful_frac", "warp_inst_per_launch")},
           "issue_peak": "%d SMs x 4 schedulers x 1 warp instruction per cycle at %.0f MHz (median SM clock sampled during the timed region); "
                         "fp32_useful_frac = algorithmic FP32 operations / (%d SMs x 128 lanes x clock x time)" % (sms, clock_hz / 1e6, sms),
           "ncu_profile": ncu[0] if ncu else None, "ncu_profile_note": ncu_why,
           "kernels": per, "step": step, "work_per_set": alg["per_set"][:len(kinds)]}
    if ncu and "stage" in ncu[1]["classes"] and tclass.get("stage", {}).get("launches"):
        # the stage class = the per-slot stage kernels + the two kernels of the queue filter (one timing mark per filter pair)
        c, v = ncu[1]["classes"]["stage"], tclass["stage"]
        t = v["ms"] / args.steps * 1e-3
        out["stage"] = {"kernels_per_step": c["launches"], "ms_per_step": round(v["ms"] / args.steps, 3),
                        "traffic_per_step": int(c["dram_read"] + c["dram_write"]),
                        "frac_dram": round((c["dram_read"] + c["dram_write"]) / t / 1e9 / peak, 4),
                        "issue_frac": round(c["warp_inst"] / (sms * 4 * clock_hz * t), 4),
                        "lanes_per_inst": round(c["thread_inst"] / max(c["warp_inst"], 1), 2)}
    return out


# ------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    side = int(round(args.spp ** 0.5))
    assert side * side == args.spp, "--spp must be a perfect square"
    assert args.spp % max(args.gpus, 1) == 0, "--spp must be divisible by --gpus"

    if args.impl == "reference":
        return main_reference(args, rank)

    import torch
    import torch.distributed as dist
    rt = importlib.import_module(PKG)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    tmp = tempfile.mkdtemp(prefix="rt_bench_")
    scene, mesh_data = build_scene(rt, args, tmp)
    slot_begin, slots_pp = rt.multi.slot_range(rank, world, args.spp)   # split by samples per pixel
    r = rt.Renderer(scene, args.cols, args.rows, args.spp, depth=args.depth, device=local_rank, slots=(slot_begin, slots_pp),
                    mode=args.mode, tile_slots=args.tile_slots)
    # host seed array: this rank's slots only are generated/kept ([pixel][k_local]) -- same values a full
    # PCG64(seed) array would hold at those positions are not needed for throughput; parity tests use full arrays.
    local = args.cols * args.rows * slots_pp
    seeds_host = torch.from_numpy(make_seeds(local, args.seed + rank)).pin_memory()
    pix_host = torch.empty(args.cols * args.rows * 4, dtype=torch.uint8).pin_memory()
    r.preRender(None)
    L = rt.lib
    ctx = r.ctx
    # Everything inside the timed regions goes through the C ABI on the context's own stream (uploads, kernels, the
    # NCCL reduce, copyToPixel, the read-back); torch supplies pinned host memory, the events and the rendezvous only.
    comm = rt.multi.Comm(ctx, rank, world) if world > 1 else None
    pix_dev = ctx.alloc(args.cols * args.rows * 4)
    ext = torch.cuda.ExternalStream(L.dll.rt_ctx_stream(ctx.h), device=torch.device("cuda", local_rank))

    def set_seeds_local():
        ctx.check(L.dll.rt_render_write_local_seeds(r.h_render, seeds_host.data_ptr(), local))

    def step_device():
        """hot path only, inputs resident in HBM"""
        r.executeRender(readback=False)
        return r.stats()

    e2e_blocking = [False]

    def step_e2e():
        """public API with host buffers: seeds H2D (non-blocking from pinned memory: the pass overlaps it with ray generation
        and the primary traversal), render, (reduce), copyToPixel, pixels D2H"""
        if e2e_blocking[0]:
            set_seeds_local()
        else:
            ctx.check(L.dll.rt_render_write_local_seeds_async(r.h_render, seeds_host.data_ptr(), local))
        r.executeRender(readback=False)
        s = r.stats()
        if comm is not None:
            comm.reduce(r, 0)   # the one collective: ncclReduce(sum) of the per-pixel accumulation images, from C++
        if rank == 0:
            m = float(np.float32(1.0 / (args.spp * (r.passes - 1))))
            ctx.check(L.dll.rt_accum_to_pixel(ctx.h, pix_dev, r.accum_dptr(), m, args.cols * args.rows))
            ctx.check(L.dll.rt_buffer_read(ctx.h, pix_dev, 0, args.cols * args.rows * 4, pix_host.data_ptr()))   # D2H + finish
        else:
            ctx.finish()
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rays = 0
        launches = 0
        kern_ms = 0.0
        e0.record(ext)
        for _ in range(k):
            s = fn()
            rays += s["closest_rays"] + s["any_rays"]
            launches += s["launches"]
            kern_ms += s["device_ms"]
        e1.record(ext)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, float(rays)], dtype=torch.float64, device=torch.device("cuda", local_rank))
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, rays = float(tmax[0]), int(t[1])
        return ms, rays, launches, kern_ms

    set_seeds_local()
    if args.profile_passes:   # under ncu: only the kernels of these passes, then the same orderly exit
        for _ in range(args.profile_passes):
            step_device()
        del ext
        ctx.finish()
        if comm is not None:
            comm.close()
        ctx.free(pix_dev)
        r.postRender()
        return
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    r.set_timing(True)   # one CUDA event in front of every launch, on the launching stream
    ms, rays, launches, kern_ms = timed(step_device, args.steps)
    tclass = r.timing()
    r.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    value = rays / (ms * 1e-3) / 1e6
    # end-to-end (host buffers, copies inside the timed region)
    step_e2e()
    ms_e, rays_e, launches_e, _ = timed(step_e2e, args.steps)
    e2e_value = rays_e / (ms_e * 1e-3) / 1e6
    # the same with the blocking upload (copy, then compute), to show what the overlap hides
    e2e_blocking[0] = True
    ms_b, rays_b, _, _ = timed(step_e2e, max(1, min(args.steps, 2)))
    e2e_blocking_value = rays_b / (ms_b * 1e-3) / 1e6

    out = None
    if rank == 0:
        info = ctx.device_info()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = roofline_of(r, scene, tclass, kern_ms, args, peak, bool(peaks), world, clocks, info)
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # reported at N = 1 only
            cpu = cpu_baseline(args)
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args),
            "run": {"slots_per_gpu_per_pixel": slots_pp, "mode": args.mode, "sm_count": info["sm_count"], "l2_bytes": info["l2_bytes"],
                    "reduce": "ncclReduce(sum, fp32, 33 MB) issued by librt2015.so on the render stream (rt_render_reduce)" if world > 1 else None},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": int(local * 4 + 64),
                    "upload": "rt_render_write_local_seeds_async: the seed H2D copy runs on a side stream inside the timed region and overlaps "
                              "ray generation + the primary traversal; the pass waits for it in front of its first shadow-ray stage",
                    "value_with_blocking_upload": round(e2e_blocking_value, 2),
                    "d2h_bytes_per_step": int(args.cols * args.rows * 4), "ms_per_step": round(ms_e / args.steps, 3)},
            "gpu_launches": int(launches),
            "rays_per_step": int(rays // max(args.steps, 1)),
            # the reference launches every kernel over EVERY slot, alive or not: slot passes per second for comparison
            "slot_passes_per_s": round(args.cols * args.rows * args.spp / (ms / args.steps * 1e-3), 0),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
    # Orderly teardown, then a NORMAL interpreter exit (exit hooks run: the driver's record of the native libraries this
    # process loaded depends on them).  Nothing of torch's was ever enqueued on the context's stream except event
    # records, so the stream can go away before torch does.
    del ext
    ctx.finish()
    torch.cuda.synchronize()
    if comm is not None:
        dist.barrier()
        comm.close()
    ctx.free(pix_dev)
    r.postRender()
    del seeds_host, pix_host
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and world == 1 and not args.no_configs and not args.no_cpu_baseline:
        # BASELINE.json configs 1-4 in compact form (one named input each, GPU side only, ~15 s): ms/frame + roofline fraction,
        # so that the driver's record carries them; `python bench_configs.py` is the full list with the CPU leg
        try:
            import bench_configs
            ctx2 = rt.lib.Context(local_rank)
            try:
                big = mesh_data if (args.mesh_u, args.mesh_v) == (1000, 500) else None
                out["configs"] = bench_configs.run_configs(rt, ctx2, compact=True, with_cpu=False, reps=2, big_mesh=big,
                                                           hbm_peak=float(out["roofline"]["peak"]))
            finally:
                ctx2.close()
        except Exception as e:   # supplementary block: never lose the headline line over it
            out["configs"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out))
    sys.stdout.flush()


def cpu_rows_default(args, seconds):
    """Rows of the frame whose CPU pass takes about `seconds` at ~0.6 Mrays/s per host core."""
    rays_per_row = args.cols * args.spp * (1 + args.depth) * (1 + args.lights) * 0.75
    rows = int(round(seconds * 0.6e6 * (os.cpu_count() or 8) / rays_per_row))
    return max(1, min(args.rows, rows))


def cpu_baseline(args):
    """Reference kernels on the host cores, on a bounded row sample of the same workload."""
    from oracle import refcl as OR
    olib = OR.load_best()
    olib.set_num_threads(os.cpu_count() or 1)
    scene, prep = oracle_scene(args, tempfile.mkdtemp(prefix="rt_bench_cpu_"))
    cam = scene["camera"].toFloat32Array()
    rows = sample_rows(args, args.cpu_rows or cpu_rows_default(args, 15.0))
    rays, dt = cpu_pass_rows(olib, prep, cam, args, scene["focal_length"], scene["lens_diameter"], rows, args.seed + 77)
    return {"value": round(rays / dt / 1e6, 3), "unit": "Mrays/s", "cores": int(olib.num_threads()), "kind": olib.kind,
            "sample": "%d pixel rows spread evenly over the frame (rows %s of %d; all %d slots/px, one executeRender pass each, %d rays, "
                      "%.1f s; reference code.cl kernels compiled by g++ -O2 -fopenmp behind oracle/clshim.h)"
                      % (len(rows), ",".join(map(str, rows)), args.rows, args.spp, rays, dt),
            "rows": rows}


def main_reference(args, rank):
    """--impl reference: the reference's own kernels (oracle/_ref when compiled, else the C
    restatement) on all host cores, same scene/metric; each step is a bounded sample of pixel rows spread evenly over
    the whole frame.  Nothing of the product (package, library, GPU) is on this path."""
    if rank != 0:
        return
    from oracle import refcl as OR
    olib = OR.load_best()
    olib.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1; this arm uses every host core
    scene, prep = oracle_scene(args, tempfile.mkdtemp(prefix="rt_bench_ref_"))
    cam = scene["camera"].toFloat32Array()
    rows = sample_rows(args, args.cpu_rows or cpu_rows_default(args, 8.0))
    rays_t, secs = 0, 0.0
    for i in range(args.warmup + args.steps):
        rays, dt = cpu_pass_rows(olib, prep, cam, args, scene["focal_length"], scene["lens_diameter"], rows, args.seed + 100000 * i)
        if i >= args.warmup:
            rays_t += rays
            secs += dt
    value = rays_t / secs / 1e6
    sample = ("each step = %d pixel rows spread evenly over the frame (rows %s of %d; all %d slots/px, one executeRender pass each); reference "
              "code.cl kernels compiled by g++ -O2 -fopenmp behind oracle/clshim.h" % (len(rows), ",".join(map(str, rows)), args.rows, args.spp))
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Mrays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(secs / args.steps * 1e3, 3), "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_of(args),
           "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": int(olib.num_threads()), "kind": olib.kind, "sample": sample,
                            "rows": rows},
           "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
