#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02e}
for n in 24 120 256; do
  AFB_CHAIN_GEOM=F timeout 120 python scratch/time_chain.py $n 2>&1 | tail -1 | cut -c1-400
done > gpurun_out/${T}_time.log 2>&1
AFB_FLOW_PROF=1 AFB_CHAIN_GEOM=F timeout 120 python scratch/time_chain.py 120 2>&1 | tail -2 | cut -c1-400 >> gpurun_out/${T}_time.log
cat gpurun_out/${T}_time.log
