#!/bin/bash
# full GPU suite (no -x: list every failure)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q ${PYTEST_ARGS} > gpurun_out/${TAG:-t}_pytest.log 2>&1
echo "pytest rc=$?"
tail -15 gpurun_out/${TAG:-t}_pytest.log
