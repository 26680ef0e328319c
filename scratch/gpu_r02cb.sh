#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02cb}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tile_lists|k_tile_mesh|k_bank_order|k_tile_stats' -c 4 -o gpurun_out/${T}_inspector python scratch/prof_inspector.py 120 1 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log; ls -la gpurun_out/${T}_inspector.ncu-rep
