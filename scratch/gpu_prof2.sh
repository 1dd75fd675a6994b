#!/bin/bash
# round-2 measurement set: GPU tests, bench (both arms), per-class ncu capture (application replay), ncu launch list,
# ncu --set full of the walker and of the stage kernel, configs 1-4 with the CPU leg.  Every ncu pass runs after its command exited 0 without ncu.
P=${1:-r2i}
mkdir -p gpurun_out
: > gpurun_out/${P}_status.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${P}_pytest.log)" >> gpurun_out/${P}_status.log
timeout 900 python bench.py > gpurun_out/${P}_bench.log 2> gpurun_out/${P}_bench.err; echo "bench rc=$?" >> gpurun_out/${P}_status.log
timeout 600 python bench.py --impl reference > gpurun_out/${P}_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/${P}_status.log
timeout 1500 ncu --replay-mode application --clock-control none --csv --log-file gpurun_out/${P}_classes.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum \
  python bench.py --profile-passes 2 > gpurun_out/${P}_ncu_classes.log 2>&1; echo "ncu classes rc=$?" >> gpurun_out/${P}_status.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv \
  python bench.py --profile-passes 2 > gpurun_out/${P}_ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/${P}_status.log
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:k_walk_pairs --launch-skip 14 --launch-count 4 \
  -o gpurun_out/${P}_walk_full -f python bench.py --profile-passes 2 > gpurun_out/${P}_ncu_walk.log 2>&1; echo "ncu walk rc=$?" >> gpurun_out/${P}_status.log
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_stage|k_filter" --launch-skip 30 --launch-count 6 \
  -o gpurun_out/${P}_stage_full -f python bench.py --profile-passes 2 > gpurun_out/${P}_ncu_stage.log 2>&1; echo "ncu stage rc=$?" >> gpurun_out/${P}_status.log
timeout 1500 python bench_configs.py --out gpurun_out/${P}_configs.json > gpurun_out/${P}_configs.log 2>&1; echo "configs rc=$?" >> gpurun_out/${P}_status.log
cat gpurun_out/${P}_status.log
