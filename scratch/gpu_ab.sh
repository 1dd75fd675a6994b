#!/bin/bash
# A/B of walker variants at the full config + GPU parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ab_smi.log 2>&1
nproc >> gpurun_out/ab_smi.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest.log
for v in macro_b3_s8 nomacro macro_b3_s16 macro_b4_s8 nomacro_b4; do
  RT2015_LIB=$PWD/scratch/ab/lib_$v.so timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ab_$v.log 2>&1
  echo "$v rc=$?" >> gpurun_out/ab_summary.log
done
