#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02j}
for v in default t512 t512b; do
  lib=""; [ $v != default ] && lib=$PWD/arcanefem_b200/variants/libafb200_$v.so
  for n in 120 256; do echo -n "$v: "; AFB200_LIB=$lib timeout 120 python scratch/time_chain.py $n 2>&1 | tail -1 | cut -c1-220; done
  echo -n "$v vec: "; AFB200_LIB=$lib timeout 200 python scratch/time_vec.py 140 1 2>&1 | tail -1 | cut -c1-220
  echo -n "$v vec: "; AFB200_LIB=$lib timeout 200 python scratch/time_vec.py 203 1 2>&1 | tail -1 | cut -c1-220
done > gpurun_out/${T}_time.log 2>&1
cat gpurun_out/${T}_time.log
