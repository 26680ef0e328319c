#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pattern_tiled -s 2 -c 2 -o gpurun_out/s15_pattern -f python scratch/prof_tiled.py > gpurun_out/s15_pattern.log 2>&1
tail -2 gpurun_out/s15_pattern.log
