#!/bin/bash
# current vector executor: ncu --set full of one steady-state launch (n=100, b=3), both layouts; event timings at n=140 / 203
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_tiled_vec -s 2 -c 1 -o gpurun_out/r02v_vec_row -f python scratch/prof_vec.py 100 row > gpurun_out/r02v_vec_row.log 2>&1
tail -2 gpurun_out/r02v_vec_row.log
timeout 300 python scratch/time_vec.py 140 1 > gpurun_out/r02v_time.log 2>&1
timeout 300 python scratch/time_vec.py 140 0 >> gpurun_out/r02v_time.log 2>&1
timeout 300 python scratch/time_vec.py 203 1 >> gpurun_out/r02v_time.log 2>&1
cat gpurun_out/r02v_time.log
