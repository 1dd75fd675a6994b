#!/bin/bash
# measurement set: GPU tests, bench (both arms), ncu launch list of the bench command, ncu --set full of walker + stage kernels
P=${1:-r1b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_status.log
timeout 900 python bench.py > gpurun_out/${P}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${P}_status.log
timeout 600 python bench.py --impl reference > gpurun_out/${P}_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/${P}_status.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${P}_ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/${P}_status.log
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:k_walk_pairs --launch-skip 24 --launch-count 4 \
   -o gpurun_out/${P}_walk_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${P}_ncu_walk.log 2>&1; echo "ncu walk rc=$?" >> gpurun_out/${P}_status.log
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:k_stage --launch-skip 24 --launch-count 4 \
   -o gpurun_out/${P}_stage_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${P}_ncu_stage.log 2>&1; echo "ncu stage rc=$?" >> gpurun_out/${P}_status.log
