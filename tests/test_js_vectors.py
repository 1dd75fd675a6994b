"""tests/golden/js_vectors.json -- the known answers `node host_node/selftest.js` checks the Node.js layer against -- must be
what the (tested) Python host produces today: the host part is regenerated here on the CPU, the grid / frame digests on the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "host_node", "make_js_vectors.py")
COMMITTED = os.path.join(ROOT, "tests", "golden", "js_vectors.json")


def _regenerate(tmp_path, *flags):
    out = str(tmp_path / "v.json")
    p = subprocess.run([sys.executable, GEN, "--out", out, *flags], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    return json.load(open(out))


def _strip(v):
    v = json.loads(json.dumps(v))
    for s in v["scenes"]:
        s.pop("grids", None)
        s.pop("frame", None)
    return v


def test_host_vectors_are_current(tmp_path):
    assert _strip(_regenerate(tmp_path)) == _strip(json.load(open(COMMITTED)))


def test_selftest_script_covers_the_vectors():
    """selftest.js cannot run here (no JavaScript engine in the image); at least every section of the vector file is consumed by it
    and every rt2015.js export it uses exists in the module's export list."""
    import re
    js = open(os.path.join(ROOT, "host_node", "selftest.js"), encoding="utf-8").read()
    mod = open(os.path.join(ROOT, "host_node", "rt2015.js"), encoding="utf-8").read()
    for key in ("cameras", "lights", "bounds", "scenes"):
        assert "V.%s" % key in js
    exports = re.search(r"module\.exports = \{(.*?)\};", mod, re.S).group(1)
    for name in set(re.findall(r"\bRT\.([A-Za-z0-9_]+)", js)):
        assert re.search(r"\b%s\b" % name, exports), "selftest.js uses RT.%s, which rt2015.js does not export" % name


@pytest.mark.gpu
def test_gpu_digests_are_current(tmp_path):
    got = _regenerate(tmp_path, "--gpu")
    want = json.load(open(COMMITTED))
    for g, w in zip(got["scenes"], want["scenes"]):
        assert "frame" in w and "grids" in w, "tests/golden/js_vectors.json lacks the GPU digests: run host_node/make_js_vectors.py --gpu"
        assert g["grids"] == w["grids"] and g["frame"] == w["frame"], g["name"]
