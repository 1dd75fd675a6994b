#!/bin/bash
mkdir -p gpurun_out
for t in 16777216 67108864 268435456 536870912; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --tile-slots $t > gpurun_out/tile_$t.log 2>&1
  echo "$t rc=$?" >> gpurun_out/tile_summary.log
done
