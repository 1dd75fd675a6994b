#!/bin/bash
mkdir -p gpurun_out
T=r2k
: > gpurun_out/${T}_summary.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)" >> gpurun_out/${T}_summary.log
timeout 600 python bench_configs.py --compact --no-cpu > gpurun_out/${T}_configs_walker.log 2>&1; echo "configs rc=$?" >> gpurun_out/${T}_summary.log
RT2015_A07_WALKER=0 timeout 600 python bench_configs.py --compact --no-cpu > gpurun_out/${T}_configs_thread.log 2>&1; echo "configs(thread) rc=$?" >> gpurun_out/${T}_summary.log
grep "synthetic mesh" gpurun_out/${T}_configs_walker.log | cut -c1-220 >> gpurun_out/${T}_summary.log
grep "synthetic mesh" gpurun_out/${T}_configs_thread.log | cut -c1-220 >> gpurun_out/${T}_summary.log
cat gpurun_out/${T}_summary.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_pytest.log | head
