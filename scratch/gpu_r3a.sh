#!/bin/bash
mkdir -p gpurun_out
T=r3a
: > gpurun_out/${T}_summary.log
run() { # name, lib, extra args
  RT2015_LIB=$2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline $3 > gpurun_out/${T}_$1.log 2> gpurun_out/${T}_$1.err
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
RT2015_LIB=$PWD/scratch/ab/lib_mark2.so timeout 900 python -m pytest tests/test_gpu_a10.py tests/test_gpu_gates.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest(mark2) rc=$? $(tail -1 gpurun_out/${T}_pytest.log)" >> gpurun_out/${T}_summary.log
run base $PWD/2015-raytracing_b200/librt2015.so
for v in mark2 mark2_MARCH_FIRST1 mark2_MARCH_FIRST4 mark2_MAX_MARCH32; do run $v $PWD/scratch/ab/lib_$v.so; done
cat gpurun_out/${T}_summary.log
