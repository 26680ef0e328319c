#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02e}
for n in 24 120 256; do
  AFB_TILED_EXEC=flow timeout 120 python scratch/time_chain.py $n 2>&1 | tail -1 | cut -c1-330
done > gpurun_out/${T}_time.log 2>&1
AFB_TILED_EXEC=flow timeout 120 python scratch/time_chain.py 2048 2 2>&1 | tail -1 | cut -c1-330 >> gpurun_out/${T}_time.log
AFB_FLOW_PROF=1 AFB_TILED_EXEC=flow timeout 120 python scratch/time_chain.py 120 2>&1 | tail -2 | cut -c1-400 >> gpurun_out/${T}_time.log
cat gpurun_out/${T}_time.log
