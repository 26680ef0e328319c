#!/bin/bash
mkdir -p gpurun_out
python scratch/prof_inspector.py 120 1 > gpurun_out/r02ae_wall.log 2>&1
python scratch/prof_inspector.py 120 1 >> gpurun_out/r02ae_wall.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ae_inspector_launches.csv python scratch/prof_inspector.py 120 1 > gpurun_out/r02ae_ncu.log 2>&1
cat gpurun_out/r02ae_wall.log
