#!/bin/bash
# launch list (per-kernel device time) of steady-state steps of scratch/prof_tiled.py
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s ${2:-30} -c ${3:-24} --csv --log-file gpurun_out/$1.csv python scratch/prof_tiled.py > gpurun_out/$1.log 2>&1
tail -1 gpurun_out/$1.log
