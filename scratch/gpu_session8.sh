#!/bin/bash
mkdir -p gpurun_out
for sk in 0 1000 2000 3000 5000; do echo "== skew $sk"; AFB_TILED_SKEW_NS=$sk timeout 300 python scratch/time_phases.py 120 2>&1 | grep tiled | cut -c1-80; done > gpurun_out/skew.log 2>&1
cat gpurun_out/skew.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pattern_tiled -s 2 -c 2 -o gpurun_out/prof_pattern python scratch/prof_tiled.py > gpurun_out/ncu_pattern.log 2>&1
