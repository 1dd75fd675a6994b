#!/bin/bash
# round 2 pass e: full GPU suite, JS vectors with GPU digests, bench with the configs block, ncu per-class capture (application replay)
mkdir -p gpurun_out
T=r2e
: > gpurun_out/${T}_status.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)" >> gpurun_out/${T}_status.log
timeout 300 python host_node/make_js_vectors.py --gpu --out gpurun_out/js_vectors.json > gpurun_out/${T}_jsvec.log 2>&1; echo "jsvec rc=$?" >> gpurun_out/${T}_status.log
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_status.log
timeout 1500 ncu --replay-mode application --clock-control none --csv --log-file gpurun_out/${T}_classes.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum \
  python bench.py --profile-passes 2 > gpurun_out/${T}_ncu_classes.log 2>&1; echo "ncu classes rc=$?" >> gpurun_out/${T}_status.log
ls -la gpurun_out/${T}_classes.csv >> gpurun_out/${T}_status.log
cat gpurun_out/${T}_status.log
