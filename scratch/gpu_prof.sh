#!/bin/bash
# round-1 measurement: bench (both arms), ncu launch list of the same command, ncu --set full of the walkers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r1_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1_status.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/r1_status.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r1_launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1_ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/r1_status.log
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:k_walk_coop --launch-skip 80 --launch-count 4 \
   -o gpurun_out/r1_walk_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rows 135 > gpurun_out/r1_ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/r1_status.log
