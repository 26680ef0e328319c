#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_rows_vec -s 2 -c 1 -o gpurun_out/r02aw_rows -f python scratch/prof_vec.py 100 row > gpurun_out/r02aw_rows.log 2>&1
tail -1 gpurun_out/r02aw_rows.log
