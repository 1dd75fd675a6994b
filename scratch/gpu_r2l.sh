#!/bin/bash
mkdir -p gpurun_out
T=r2l
: > gpurun_out/${T}_summary.log
timeout 900 python -m pytest tests/test_gpu_gates.py tests/test_gpu_golden.py -m gpu -q -k "a07 or mesh or molecule" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)" >> gpurun_out/${T}_summary.log
timeout 600 python bench_configs.py --compact --no-cpu > gpurun_out/${T}_configs_walker.log 2>&1; echo "configs rc=$?" >> gpurun_out/${T}_summary.log
RT2015_A07_WALKER=0 timeout 600 python bench_configs.py --compact --no-cpu > gpurun_out/${T}_configs_thread.log 2>&1; echo "configs(thread) rc=$?" >> gpurun_out/${T}_summary.log
grep "synthetic mesh" gpurun_out/${T}_configs_walker.log | cut -c1-200 >> gpurun_out/${T}_summary.log
grep "synthetic mesh" gpurun_out/${T}_configs_thread.log | cut -c1-200 >> gpurun_out/${T}_summary.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_a07_launches.csv --kernel-name regex:"a07|walk_pairs" python bench_configs.py --compact --no-cpu > /dev/null 2>&1
grep -c . gpurun_out/${T}_a07_launches.csv >> gpurun_out/${T}_summary.log
cat gpurun_out/${T}_summary.log
