#!/bin/bash
# lean final set after a kernel-source change: GPU tests, quick bench, per-class ncu capture, launch list
P=${1:-r2g}; export PTAG=$P
mkdir -p gpurun_out
: > gpurun_out/${P}_status.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${P}_pytest.log)" >> gpurun_out/${P}_status.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${P}_quick.log 2> gpurun_out/${P}_quick.err; echo "quick rc=$?" >> gpurun_out/${P}_status.log
timeout 1500 ncu --replay-mode application --clock-control none --csv --log-file gpurun_out/${P}_classes.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum \
  python bench.py --profile-passes 2 > gpurun_out/${P}_ncu_classes.log 2>&1; echo "ncu classes rc=$?" >> gpurun_out/${P}_status.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv \
  python bench.py --profile-passes 2 > gpurun_out/${P}_ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/${P}_status.log
cat gpurun_out/${P}_status.log
tail -c 300 gpurun_out/${P}_quick.log
timeout 900 python bench.py > gpurun_out/${P}_bench.log 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${P}_ref.log 2>&1; echo "ref rc=$?"
