#!/bin/bash
mkdir -p gpurun_out
T=${1:-tests}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -12 gpurun_out/${T}_pytest.log
