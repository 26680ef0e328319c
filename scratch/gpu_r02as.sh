#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02as}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_inspector_launches.csv python scratch/prof_inspector.py 120 1 > gpurun_out/${T}_ncu.log 2>&1
python scratch/prof_inspector.py 256 1 > gpurun_out/${T}_wall.log 2>&1; cat gpurun_out/${T}_wall.log
