#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tile_lists -c 1 -o gpurun_out/r02ah_lists -f python scratch/prof_inspector.py 120 1 > gpurun_out/r02ah_lists.log 2>&1
tail -1 gpurun_out/r02ah_lists.log
