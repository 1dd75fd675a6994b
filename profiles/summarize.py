#!/usr/bin/env python3
"""Turns the scratch outputs of scratch/gpu_prof.sh (gpurun_out/<P>_*) into the tracked summaries under
profiles/:  <P>_bench.json, <P>_bench_reference.json, <P>_launches.csv.gz, <P>_launches_summary.md,
<P>_walk_full_summary.md, <P>_stage_full_summary.md.      python profiles/summarize.py r1b
Needs `ncu` (to read the .ncu-rep files)."""
import collections
import csv
import gzip
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")
P = sys.argv[1] if len(sys.argv) > 1 else "r1b"

KEEP = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']


def short(name):
    name = re.sub(r'<unnamed>::|\(int\)|\(bool\)', '', name)
    return name.split('(')[0].replace('void ', '').strip()


def launches():
    rows, hdr = [], None
    with open(os.path.join(G, P + "_launches.csv")) as f:
        for line in f:
            if line.startswith('"ID"'):
                hdr = next(csv.reader([line]))
                break
        for r in csv.reader(f):
            if len(r) == len(hdr):
                rows.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r['Metric Value']) / 1e6
        a = agg.setdefault(short(r['Kernel Name']), [0, 0.0, 1e9, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    bench = json.load(open(os.path.join(OUT, P + "_bench.json")))
    cls = bench["roofline"]["step"]["class_ms_per_step"]
    ctot = sum(cls.values())
    with open(os.path.join(OUT, P + "_launches_summary.md"), "w") as f:
        f.write("# %s -- ncu launch list of the bench command\n\n" % P)
        cmd = ("python bench.py --steps 1 --warmup 1 --no-cpu-baseline`\n(scene build + warm-up pass + timed pass + the passes of the e2e/profile legs, " if P.startswith("r1")
               else "python bench.py --profile-passes 2`\n(scene build + two passes, ")
        f.write(("Command (B200, 1 GPU): `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file %s_launches.csv " + cmd +
                 "full config: 1920x1080, 256 slots/px, 1 M-triangle mesh, nslabs 128; %d launches captured).\n"
                 "Per-launch times under ncu are serialised and cold-cache; what must agree with bench.py's CUDA-event timing is each kernel's "
                 "SHARE of the step.\n\n") % (P, len(rows)))
        step = lambda k: k.startswith(("k_stage", "k_walk", "k_filter", "f_sumSlots"))   # the kernels of a pass (k_pathMega<1> = the instrumented work-profile pass)
        stot = sum(a[1] for k, a in agg.items() if step(k))
        f.write("| kernel | launches | total ms | avg ms | min ms | max ms | share of all | share of the pass kernels |\n|---|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.2f | %.4f | %.4f | %.4f | %.1f %% | %s |\n" % (k, a[0], a[1], a[1] / a[0], a[2], a[3], 100 * a[1] / tot,
                                                                                 "%.1f %%" % (100 * a[1] / stot) if step(k) else "-"))
        f.write("\nbench.py, same build, CUDA events on the launching stream, no profiler (ms per step): " +
                ", ".join("%s %.1f (%.1f %%)" % (k, v, 100 * v / ctot) for k, v in cls.items()) + ".\n")
    with open(os.path.join(G, P + "_launches.csv"), 'rb') as s, gzip.open(os.path.join(OUT, P + "_launches.csv.gz"), 'wb') as d:
        shutil.copyfileobj(s, d)


def full(tag, title, kernel_filter):
    rep = os.path.join(G, "%s_%s_full.ncu-rep" % (P, tag))
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = io.StringIO()
    out.write("# %s -- `ncu --set full` of %s\n\n" % (P, title))
    if P.startswith("r1"):
        out.write("Command (B200, 1 GPU): `ncu --set full --clock-control none --import-source on --kernel-name regex:%s --launch-skip 24 "
                  "--launch-count 4 -o %s_%s_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline` (full config; the frame is two tiles, 412 721 664 + 118 119 936 "
                  "slots, a launch serves one tile; the captured launches belong to the small tile).\n\n" % (kernel_filter, P, tag))
    else:
        out.write("Command (B200, 1 GPU): `ncu --set full --clock-control none --import-source on --kernel-name regex:\"%s\" --launch-skip ... "
                  "--launch-count ... -o %s_%s_full python bench.py --profile-passes 2` (scratch/gpu_prof2.sh; full config, two passes; the frame is two "
                  "wavefront tiles and a launch serves one tile; the window falls into the second pass).\n\n" % (kernel_filter, P, tag))
    n = len(rows) - 2
    i_n = hdr.index('Kernel Name')
    out.write("| metric | unit | " + " | ".join(short(r[i_n]) for r in rows[2:]) + " |\n|---|---|" + "---|" * n + "\n")
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            out.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in rows[2:])))
    i_r, i_w = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    mul = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1, 'Tbyte': 1e12}
    tr = collections.OrderedDict()
    for r in rows[2:]:
        tr.setdefault(short(r[i_n]), []).append(float(r[i_r].replace(',', '')) * mul[units[i_r]] + float(r[i_w].replace(',', '')) * mul[units[i_w]])
    out.write("\nDRAM traffic per launch (read + write): " + "; ".join("`%s`: mean %.2f GB over %d launches" % (k, sum(v) / len(v) / 1e9, len(v))
                                                                        for k, v in tr.items()) + "\n")
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    agg = collections.defaultdict(lambda: [0, 0, 0, ""])
    fn = fil = None
    h2 = None
    total = 0
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            fil = r[1].split('/')[-1]; continue
        if r[0] == "Function Name":
            fn = r[1]; continue
        if r[0] == "Line No":
            h2 = r; iI = h2.index("Instructions Executed"); iT = h2.index("Thread Instructions Executed"); iS = h2.index("# Samples"); continue
        if r[0] == "" or h2 is None or len(r) != len(h2):
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        k = (fil, line)
        agg[k][0] += int(r[iI]); agg[k][1] += int(r[iT]); agg[k][2] += int(r[iS]); agg[k][3] = r[1].strip()[:105]
        total += int(r[iI])
    out.write("\n## Hot source lines (warp instructions executed, share, threads per instruction, stall samples) -- all captured launches\n\n```\n")
    out.write("total warp instructions: %d\n" % total)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:40]:
        out.write("%12d %5.1f%% %5.1f %8d  %s:%d  %s\n" % (v[0], 100.0 * v[0] / max(total, 1), v[1] / max(v[0], 1), v[2], k[0], k[1], v[3]))
    out.write("```\n")
    open(os.path.join(OUT, "%s_%s_full_summary.md" % (P, tag)), "w").write(out.getvalue())
    return tr


def bench_lines():
    for src, dst in ((P + '_bench.log', P + '_bench.json'), (P + '_ref.log', P + '_bench_reference.json')):
        for l in open(os.path.join(G, src)):
            if l.startswith('{'):
                json.dump(json.loads(l), open(os.path.join(OUT, dst), 'w'), indent=1)


if __name__ == "__main__":
    bench_lines()
    launches()
    print(full("walk", "the dominant kernel (pair-list queue walker over the 1 M-triangle mesh grid)", "k_walk_pairs"))
    print(full("stage", "the per-slot stage kernels (ray generation, 1-cell sets, shading, queue push) and the queue filter",
               "k_stage" if P.startswith("r1") else "k_stage|k_filter"))
