#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02h}
for v in default r768a r768b r1024; do
  lib=""; [ $v != default ] && lib=$PWD/scratch/libs/libafb200_$v.so
  for n in 120 256; do
    echo -n "$v: "; AFB200_LIB=$lib AFB_CHAIN_GEOM=F timeout 120 python scratch/time_chain.py $n 2>&1 | tail -1 | cut -c1-200
  done
done > gpurun_out/${T}_time.log 2>&1
cat gpurun_out/${T}_time.log
