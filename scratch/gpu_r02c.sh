#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02c}
AFB_CHAIN_GEOM=B timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_chain -s 3 -c 1 -o gpurun_out/${T}_chainB -f python scratch/time_chain.py 120 > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
AFB_CHAIN_GEOM=A timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_chain -s 3 -c 1 -o gpurun_out/${T}_chainA -f python scratch/time_chain.py 120 >> gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
