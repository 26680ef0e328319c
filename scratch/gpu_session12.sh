#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_tiled -s 2 -c 1 -o gpurun_out/s12_tiled -f python scratch/prof_tiled.py > gpurun_out/s12_ncu.log 2>&1
tail -3 gpurun_out/s12_ncu.log; ls -la gpurun_out/
