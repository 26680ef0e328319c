#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02bx}
python scratch/prof_inspector3.py 120 120 256 120 2> gpurun_out/${T}_inspector_trace.txt; grep "first tiled" gpurun_out/${T}_inspector_trace.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
