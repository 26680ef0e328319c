#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02g}
AFB_CHAIN_GEOM=F timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_assemble_flow -s 3 -c 1 -o gpurun_out/${T}_flow -f python scratch/time_chain.py 120 > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
