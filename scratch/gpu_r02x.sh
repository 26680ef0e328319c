#!/bin/bash
mkdir -p gpurun_out
AFB_VEC_EXEC=rows timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_rows_vec -s 2 -c 1 -o gpurun_out/r02x_rows -f python scratch/prof_vec.py 100 row > gpurun_out/r02x_rows.log 2>&1
tail -2 gpurun_out/r02x_rows.log
