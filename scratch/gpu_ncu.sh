#!/bin/bash
# full ncu capture (with source) of one steady-state launch of a kernel: gpu_ncu.sh <regex> <outname> [script args]
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 -s ${4:-2} -c 1 -o gpurun_out/$2 -f python ${3:-scratch/prof_tiled.py} > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log
