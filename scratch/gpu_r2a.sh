#!/bin/bash
# round 2, first GPU pass: full GPU suite, bench (both arms), exit status of the bench process
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_status.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_status.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_status.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.log 2>&1
echo "ref rc=$?" >> gpurun_out/r2a_status.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2a_status.log
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_status.log
