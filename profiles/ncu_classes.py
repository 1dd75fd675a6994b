#!/usr/bin/env python3
"""ncu per-launch metrics of ONE pass of the bench workload -> profiles/<tag>_ncu_classes.json, the file bench.py reads
for `roofline.traffic` (measured DRAM bytes per launch), `frac_dram` and the issue roofline (warp / thread instructions
per launch).  The JSON is keyed by a hash of the kernel sources it was captured from; bench.py refuses it when the
sources have changed since (a stale profile must not be scaled into a new build's numbers).

    # on the GPU box (scratch/gpu_prof2.sh): two passes of the full config, every launch, application replay
    ncu --replay-mode application --clock-control none --csv --log-file gpurun_out/<tag>_classes.csv \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum \
        python bench.py --profile-passes 2
    # here
    python profiles/ncu_classes.py gpurun_out/<tag>_classes.csv profiles/<tag>_ncu_classes.json

The LAST pass's launches are kept (the first is the warm-up).  Classes follow bench.py's timing classes."""
import collections
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["2015-raytracing_b200/csrc/rt_wavefront.cu", "2015-raytracing_b200/csrc/rt_device.cuh", "2015-raytracing_b200/csrc/rt_frame.cu",
           "2015-raytracing_b200/csrc/Makefile"]
METRICS = {"gpu__time_duration.sum": "time_ns", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
           "smsp__inst_executed.sum": "warp_inst", "smsp__thread_inst_executed.sum": "thread_inst"}
UNIT = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9, "second": 1e9,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "inst": 1.0, "": 1.0}


def source_sha(root=ROOT):
    h = hashlib.sha256()
    for rel in SOURCES:
        with open(os.path.join(root, rel), "rb") as f:
            h.update(rel.encode() + b"\0" + f.read() + b"\0")
    return h.hexdigest()[:16]


def classify(kernel):
    k = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", kernel)
    m = re.search(r"k_walk_pairs<\(?(?:int\))?(\d),\s*\(?(?:bool\))?(\d|true|false)\s*[,>]", k)
    if m:
        prim = "sphere" if m.group(1) == "0" else "triangle"
        anyh = m.group(2) in ("1", "true")
        return "walk_%s_%s" % (prim, "any" if anyh else "closest")
    if "k_stage<" in k or "k_filter" in k:   # the queue filter is charged to the stage class by bench.py's timing too
        return "stage"
    if "f_sumSlots" in k or "f_accumToPixel" in k:
        return "sum_copy"
    if "k_pathMega" in k:
        return "megakernel"
    return None


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows, hdr = [], None
    with open(src) as f:
        for line in f:
            if line.startswith('"ID"'):
                hdr = next(csv.reader([line]))
                break
        for r in csv.reader(f):
            if hdr and len(r) == len(hdr):
                rows.append(dict(zip(hdr, r)))
    launches = collections.OrderedDict()   # ID -> {kernel, metrics}
    for r in rows:
        name = METRICS.get(r["Metric Name"])
        if not name:
            continue
        L = launches.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"]})
        L[name] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    seq = [(i, L, classify(L["kernel"])) for i, L in sorted(launches.items())]
    seq = [(i, L, c) for i, L, c in seq if c]
    # a pass ends with f_sumSlots: keep the launches after the second-to-last one
    ends = [n for n, (_, _, c) in enumerate(seq) if c == "sum_copy"]
    if len(ends) >= 2:
        seq = seq[ends[-2] + 1:ends[-1] + 1]
    classes = collections.OrderedDict()
    for _, L, c in seq:
        a = classes.setdefault(c, collections.OrderedDict([("launches", 0)] + [(v, 0.0) for v in METRICS.values()]))
        a["launches"] += 1
        for v in METRICS.values():
            a[v] += L.get(v, 0.0)
    out = {"source_sha": source_sha(), "sources": SOURCES, "captured_launches": len(seq),
           "command": "ncu --replay-mode application --clock-control none --metrics %s python bench.py --profile-passes 2 (full config, N = 1; "
                      "the launches of the last pass)" % ",".join(METRICS),
           "classes": classes}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    for c, a in classes.items():
        print("%-24s launches %3d  time %9.3f ms  dram %8.2f GB  warp inst %.3e  lanes/inst %.1f" % (
            c, a["launches"], a["time_ns"] / 1e6, (a["dram_read"] + a["dram_write"]) / 1e9, a["warp_inst"],
            a["thread_inst"] / a["warp_inst"] if a["warp_inst"] else 0))


if __name__ == "__main__":
    main()
