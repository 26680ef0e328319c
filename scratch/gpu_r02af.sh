#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tile_lists -c 1 -o gpurun_out/r02af_lists -f python scratch/prof_inspector.py 120 1 > gpurun_out/r02af_lists.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tile_mesh -c 1 -o gpurun_out/r02af_mesh -f python scratch/prof_inspector.py 120 1 > gpurun_out/r02af_mesh.log 2>&1
tail -1 gpurun_out/r02af_mesh.log
