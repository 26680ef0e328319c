#!/bin/bash
mkdir -p gpurun_out
AFB_TILED_EXEC=pipe timeout 300 python scratch/time_phases.py 120 2>&1 | grep tiled | tee gpurun_out/q_time_pipe.log
timeout 300 python scratch/time_phases.py 120 2>&1 | grep tiled | tee gpurun_out/q_time.log
