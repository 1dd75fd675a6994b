"""The committed ncu evidence must belong to the kernel sources in the tree: bench.py takes `roofline.traffic`, `frac_dram`
and the issue roofline from profiles/*_ncu_classes.json and refuses a file whose source hash differs (it then prints
`ncu_profile_note: stale`).  This test makes an edit of the walker / stage sources without a fresh capture visible here,
on the CPU, instead of on the bench line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _profiles():
    pdir = os.path.join(ROOT, "profiles")
    out = []
    for fn in sorted(os.listdir(pdir)):
        if fn.endswith("_ncu_classes.json"):
            out.append((fn, json.load(open(os.path.join(pdir, fn)))))
    return out


def test_a_class_profile_of_the_current_kernel_sources_is_committed():
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import ncu_classes
    sha = ncu_classes.source_sha(ROOT)
    fresh = [fn for fn, d in _profiles() if d.get("source_sha") == sha]
    assert fresh, "no profiles/*_ncu_classes.json was captured from the current kernel sources (sha %s): re-run " \
                  "scratch/gpu_final2.sh and profiles/ncu_classes.py" % sha


def test_class_profiles_are_whole_passes():
    # one pass of the default workload: 12 + 12 walks over the mesh grid, the stage class (stages + queue filter), one sum
    for fn, d in _profiles():
        c = d["classes"]
        assert c["walk_triangle_any"]["launches"] == 12 and c["walk_triangle_closest"]["launches"] == 12, fn
        assert c["sum_copy"]["launches"] == 1, fn
        for k, v in c.items():
            assert v["time_ns"] > 0 and v["warp_inst"] > 0 and v["thread_inst"] >= v["warp_inst"], (fn, k)
            assert v["thread_inst"] <= 32 * v["warp_inst"], (fn, k)


def test_bench_accepts_the_committed_profile():
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    cols, rows, spp, depth, mesh_u, mesh_v, nslabs, lights = bench.NCU_WORKLOAD
    args = argparse.Namespace(cols=cols, rows=rows, spp=spp, depth=depth, mesh_u=mesh_u, mesh_v=mesh_v, nslabs=nslabs, lights=lights, mode=0)
    best, note = bench.load_ncu_classes(args, 1)
    assert note is None and best is not None, note
    best, note = bench.load_ncu_classes(args, 8)
    assert best is None and "N = 1" in note
    args.spp = 4
    best, note = bench.load_ncu_classes(args, 1)
    assert best is None and "non-default" in note


def _fmt(v):
    s = "%d" % round(v)
    return s[:-3] + " " + s[-3:] if len(s) > 3 else s   # "8 705" as the documents write it


def test_documents_quote_the_committed_bench_lines():
    """README.md and DESIGN.md quote the headline numbers; they must be the ones of the newest committed bench lines."""
    pdir = os.path.join(ROOT, "profiles")
    sha_of = {fn: d.get("source_sha") for fn, d in _profiles()}
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import ncu_classes
    tag = [fn.split("_")[0] for fn, sha in sha_of.items() if sha == ncu_classes.source_sha(ROOT)][-1]
    line = json.load(open(os.path.join(pdir, tag + "_bench.json")))
    assert line["n_gpus"] == 1 and line["roofline"]["ncu_profile"] == tag + "_ncu_classes.json"
    want = [_fmt(line["value"]), _fmt(line["e2e"]["value"])]
    n8 = os.path.join(pdir, tag + "_bench_n8.json")
    if os.path.exists(n8):
        want.append(_fmt(json.load(open(n8))["value"]))
    for doc in ("README.md", "DESIGN.md"):
        text = open(os.path.join(ROOT, doc), encoding="utf-8").read().replace(" ", " ").replace(" ", " ")
        for w in want:
            assert w in text, "%s does not quote %s (profiles/%s_bench*.json)" % (doc, w, tag)
