#!/bin/bash
# round 2 A/B #2: split vs merged lights at full and at N=8-sized load, cull-load pipelining variants; GPU suite in both light modes
mkdir -p gpurun_out
T=${TAG:-ab2}
: > gpurun_out/${T}_summary.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest merged rc=$? $(tail -1 gpurun_out/${T}_pytest.log)" >> gpurun_out/${T}_summary.log
RT2015_SPLIT_LIGHTS=1 timeout 1500 python -m pytest tests/test_gpu_a10.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/${T}_pytest_split.log 2>&1
echo "pytest split rc=$? $(tail -1 gpurun_out/${T}_pytest_split.log)" >> gpurun_out/${T}_summary.log
run() { # name, lib, extra args
  RT2015_LIB=$2 timeout 600 python bench.py --steps ${STEPS:-2} --warmup ${WARMUP:-1} --no-cpu-baseline $3 > gpurun_out/${T}_$1.log 2> gpurun_out/${T}_$1.err
  rc=$?
  python - "$1" "$rc" gpurun_out/${T}_$1.log >> gpurun_out/${T}_summary.log <<'PY'
import json, sys
name, rc, path = sys.argv[1:4]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    c = d["roofline"]["step"]["class_ms_per_step"]
    print("%-14s rc=%s value=%8.1f e2e=%8.1f ms=%8.2f  %s" % (name, rc, d["value"], d["e2e"]["value"], d["ms_per_step"], " ".join("%s=%.1f" % (k, v) for k, v in c.items())))
except Exception as e:
    print("%-14s rc=%s FAILED %r" % (name, rc, e))
PY
}
B=$PWD/2015-raytracing_b200/librt2015.so
run merged $B
RT2015_SPLIT_LIGHTS=1 run split $B
run merged36 $B "--spp 36"
RT2015_SPLIT_LIGHTS=1 run split36 $B "--spp 36"
for f in scratch/ab/lib_*.so; do
  v=$(basename $f .so); v=${v#lib_}
  run $v $PWD/$f
done
cat gpurun_out/${T}_summary.log
