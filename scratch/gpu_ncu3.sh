#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scratch/prof_vec.py 100 2>&1 | tail -1
timeout 120 python scratch/prof_vec.py 100 row 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_tiled_vec -s 2 -c 1 -o gpurun_out/s17_vec -f python scratch/prof_vec.py 100 > gpurun_out/s17_vec.log 2>&1
tail -2 gpurun_out/s17_vec.log
