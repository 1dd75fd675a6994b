#!/bin/bash
# wall skip + fine distance field + chunk0: correctness + A/B
mkdir -p gpurun_out
T=r2h
: > gpurun_out/${T}_summary.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)" >> gpurun_out/${T}_summary.log
run() { # name, extra args
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline $2 > gpurun_out/${T}_$1.log 2> gpurun_out/${T}_$1.err
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
run all
RT2015_NO_WALL_SKIP=1 run nowall
RT2015_NO_SKIP=1 run noskip
RT2015_NO_SKIP=1 RT2015_NO_WALL_SKIP=1 run neither
run all36 "--spp 36"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --profile-passes 1 > gpurun_out/${T}_ncu_list.log 2>&1
cat gpurun_out/${T}_summary.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_pytest.log | head
