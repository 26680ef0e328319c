#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 300 python scratch/time_phases.py 120 100 > gpurun_out/phases.log 2>&1; cat gpurun_out/phases.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiled -s 2 -c 1 -o gpurun_out/prof_tiled python scratch/prof_tiled.py > gpurun_out/ncu_tiled.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pattern_tiled -s 2 -c 2 -o gpurun_out/prof_pattern python scratch/prof_tiled.py > gpurun_out/ncu_pattern.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches.csv python scratch/prof_tiled.py > gpurun_out/ncu_l.log 2>&1
