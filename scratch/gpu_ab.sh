#!/bin/bash
# A/B of library variants (scratch/ab/lib_*.so, same ABI) on the full bench config
mkdir -p gpurun_out
for f in scratch/ab/lib_*.so; do
  v=$(basename $f .so); v=${v#lib_}
  RT2015_LIB=$PWD/$f timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ab_$v.log 2>&1
  echo "$v rc=$?" >> gpurun_out/ab_summary.log
done
