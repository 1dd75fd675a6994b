#!/bin/bash
P=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${P}_bench.log 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${P}_ref.log 2>&1; echo "ref rc=$?"
tail -c 600 gpurun_out/${P}_bench.log
