#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp.log
for f in arcanefem_b200/variants/libafb200_*.so; do
  v=$(basename $f .so); v=${v#libafb200_}
  echo "== $v $(AFB200_LIB=$PWD/$f timeout 120 python scratch/time_phases.py 120 2>&1 | grep 'tiled/conn' | cut -c1-80)" >> gpurun_out/exp.log
done
echo "== base $(timeout 120 python scratch/time_phases.py 120 2>&1 | grep 'tiled/conn' | cut -c1-80)" >> gpurun_out/exp.log
cat gpurun_out/exp.log
